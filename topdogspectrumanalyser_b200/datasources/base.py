"""The boundary class. Inside the reference application (its root on ``sys.path``) this IS the
reference's ``datasources.base.SampleDataSource``; stand-alone (tests, GPU box) it is a mirror with
the same members (datasources/base.py:43-169): ``_averager``, ``_last_raw_samples`` under
``_raw_lock``, ``last_data_time``, abstract start/stop/get_power_levels/sample_count/
update_frequency/update_centre_frequency, and set_psd_mode/set_averaging/reset_averaging.
"""
from __future__ import annotations

import threading
import time
from abc import ABC, abstractmethod
from typing import Optional, Tuple

import numpy as np

try:                                     # running inside the reference app
    from datasources.base import SampleDataSource as _RefSampleDataSource   # type: ignore
    from datasources.base import SweepDataSource as _RefSweepDataSource     # type: ignore
    IN_REFERENCE_APP = True
except Exception:                        # stand-alone
    _RefSampleDataSource = None
    _RefSweepDataSource = None
    IN_REFERENCE_APP = False


if _RefSweepDataSource is not None:
    SweepDataSource = _RefSweepDataSource
else:
    class SweepDataSource(ABC):
        """Mirror of datasources/base.py:15-39."""

        @abstractmethod
        def start(self, frequency=None):
            ...

        @abstractmethod
        def stop(self):
            ...

        @abstractmethod
        def get_data(self):
            ...


class AveragerSettings:
    """Holds what ``TraceAverager.set_mode/reset`` configure (utils/signal_processing.py:19-33,63-73).

    The arithmetic of ``process`` runs on the device (tdsa_psd_db_avg_hold); this object only
    carries mode/n and forwards resets, so callers that poke ``source._averager`` keep working.
    """

    def __init__(self, on_change=None):
        self._mode, self._n = "off", 1
        self._on_change = on_change

    def set_mode(self, mode: str, n: int) -> None:
        self._mode, self._n = mode, max(1, n)
        self.reset()

    def reset(self) -> None:
        if self._on_change is not None:
            self._on_change()

    def process(self, linear_power):
        raise RuntimeError("averaging runs on the device (SpectrumPlan.psd_db_avg_hold); no host path")

    @property
    def is_active(self) -> bool:
        return self._mode != "off" and self._n > 1

    @property
    def mode(self) -> str:
        return self._mode

    @property
    def n(self) -> int:
        return self._n


if _RefSampleDataSource is not None:
    SampleDataSource = _RefSampleDataSource
else:
    class SampleDataSource(ABC):
        """Mirror of datasources/base.py:43-169 (same names, argument meaning and defaults)."""

        def __init__(self, sample_rate: Optional[int] = None, centre_freq: Optional[int] = None):
            self.sample_rate = sample_rate
            self.centre_freq = centre_freq
            self._averager = AveragerSettings()
            self._last_raw_samples: Optional[np.ndarray] = None
            self.last_data_time: float = 0.0
            self._raw_lock = threading.Lock()

        @abstractmethod
        def start(self, frequency=None):
            ...

        @abstractmethod
        def stop(self):
            ...

        @abstractmethod
        def get_power_levels(self) -> Tuple[np.ndarray, np.ndarray]:
            ...

        @property
        @abstractmethod
        def sample_count(self) -> int:
            ...

        @sample_count.setter
        @abstractmethod
        def sample_count(self, value: int):
            ...

        @abstractmethod
        def update_frequency(self, sample_rate: float, centre_freq: float):
            ...

        @abstractmethod
        def update_centre_frequency(self, centre_freq: float):
            ...

        def get_raw_samples(self) -> Optional[np.ndarray]:
            with self._raw_lock:
                return self._last_raw_samples

        def read_samples_only(self) -> Optional[np.ndarray]:
            return None

        def _store_raw(self, samples: np.ndarray) -> None:
            with self._raw_lock:
                self._last_raw_samples = samples
            self.last_data_time = time.monotonic()

        def set_psd_mode(self, enabled: bool):
            pass

        def set_averaging(self, mode: str, n: int) -> None:
            self._averager.set_mode(mode, n)

        def reset_averaging(self) -> None:
            self._averager.reset()
