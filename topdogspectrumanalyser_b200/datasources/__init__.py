from .base import SampleDataSource, AveragerSettings, IN_REFERENCE_APP
from .b200_samples import B200SampleDataSource, HackrfChunkFeed, ReplayFeed, SyntheticIQFeed, register_with_source_manager

__all__ = ["SampleDataSource", "AveragerSettings", "IN_REFERENCE_APP", "B200SampleDataSource",
           "SyntheticIQFeed", "HackrfChunkFeed", "ReplayFeed", "register_with_source_manager"]
