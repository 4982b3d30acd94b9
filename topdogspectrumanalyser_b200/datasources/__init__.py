from .base import SampleDataSource, SweepDataSource, AveragerSettings, IN_REFERENCE_APP
from .feeds import ChunkRingFeed, HackrfDeviceFeed, ReplayFeed, SyntheticIQFeed, open_rtlsdr
from .b200_samples import (B200HackrfSamples, B200RtlSamples, B200SampleDataSource, HackrfChunkFeed, install_backend,
                           uninstall_backend)

from .b200_sweep import B200SweepDataSource, SyntheticTunerFeed

__all__ = ["SweepDataSource", "B200SweepDataSource", "SyntheticTunerFeed", "SampleDataSource", "AveragerSettings", "IN_REFERENCE_APP", "B200SampleDataSource", "B200RtlSamples",
           "B200HackrfSamples", "SyntheticIQFeed", "ChunkRingFeed", "HackrfChunkFeed", "HackrfDeviceFeed", "ReplayFeed",
           "open_rtlsdr", "install_backend", "uninstall_backend"]
