"""Device-resident counterpart of ``DataProcessor._process_sample_data`` (core/display_data_processor.py:153-183).

The reference fetches a frame, then on the host: cal offset (:317-327) -> tare (:329-369) -> live trace ->
max/min hold (:371-395) -> top-5 peak list (:407-471).  Here the frame never leaves the GPU between the fused FFT
kernel and those stages; attributes carry the reference's names (``live_power_levels``, ``max_power_levels``,
``min_power_levels``, ``frequency_bins``, ``tare_active``) and hand out host arrays on demand.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from .analytics import top_peaks
from .engine import TraceState, trace_update


class B200FramePipeline:
    def __init__(self, source, cal_offset_db: float = 0.0, peak_list: bool = False, peak_excursion: float = 10.0):
        self.source = source
        self.cal_offset_db = float(cal_offset_db)
        self._max_on = False                          # dm.max_peak_search_enabled
        self._min_on = False                          # mw.min_hold_enabled
        self.peak_list_enabled = peak_list
        self.peak_excursion = float(peak_excursion)   # mw.peak_excursion (display_data_processor.py:418)
        self.frequency_bins: Optional[np.ndarray] = None
        self.peaks: List[Tuple[float, float]] = []
        self._state: Optional[TraceState] = None
        self._live: Optional[torch.Tensor] = None

    # ---- controls (display_manager.py:139-198,  _tare_action) -------------------------------------
    def _ensure_state(self, width: int, device) -> TraceState:
        if self._state is None or self._state.width != width:        # shape change drops holds and tare (:375-377,361-364)
            self._state = TraceState(width, device)
        # a disabled hold keeps its row until the shape changes (:374-378); enabling starts afresh (the setters below)
        self._state.max_hold_enabled = self._max_on
        self._state.min_hold_enabled = self._min_on
        return self._state

    @property
    def max_peak_search_enabled(self) -> bool:
        return self._max_on

    @max_peak_search_enabled.setter
    def max_peak_search_enabled(self, on: bool) -> None:
        """display_manager.py:139-157: every enable starts from an empty hold."""
        on = bool(on)
        if on and not self._max_on and self._state is not None:
            self._state.clear_max_hold()
        self._max_on = on

    @property
    def min_hold_enabled(self) -> bool:
        return self._min_on

    @min_hold_enabled.setter
    def min_hold_enabled(self, on: bool) -> None:
        """display_manager.py:166-176: the buffer is reset on every toggle."""
        on = bool(on)
        if on != self._min_on and self._state is not None:
            self._state.clear_min_hold()
        self._min_on = on

    def start_tare(self) -> None:
        if self._state is not None:
            self._state.start_tare()
        else:
            self._pending_tare = True

    def clear_tare(self) -> None:
        if self._state is not None:
            self._state.clear_tare()

    def clear_hold(self) -> None:
        if self._state is not None:
            self._state.clear_holds()

    @property
    def tare_active(self) -> bool:
        return self._state is not None and self._state.tare_active

    # ---- one tick ------------------------------------------------------------------------------------
    def update_data(self) -> bool:
        """One frame through source -> cal -> tare -> holds -> peaks. Returns False if no new frame arrived."""
        row, bins = self.source.get_power_levels_device()
        if row is None:
            return False
        st = self._ensure_state(row.shape[1], row.device)
        if getattr(self, "_pending_tare", False):
            st.start_tare()
            self._pending_tare = False
        self.frequency_bins = bins
        self._live = trace_update(row, st, self.cal_offset_db)
        if self.peak_list_enabled:
            # display_data_processor.py:417-418: separation scales with the trace width
            self.peaks = top_peaks(bins, self._live[0], min_sep_bins=max(10, len(bins) // 50),
                                   min_excursion_db=self.peak_excursion)
        return True

    # ---- what the widgets read ---------------------------------------------------------------------------
    @property
    def live_power_levels(self) -> Optional[np.ndarray]:
        return None if self._live is None else self._live[0].cpu().numpy()

    @property
    def max_power_levels(self) -> Optional[np.ndarray]:
        st = self._state
        return st.max_hold.cpu().numpy() if st is not None and st.valid[0] else None

    @property
    def min_power_levels(self) -> Optional[np.ndarray]:
        st = self._state
        return st.min_hold.cpu().numpy() if st is not None and st.valid[1] else None

    @property
    def baseline_power_levels(self) -> Optional[np.ndarray]:
        st = self._state
        return st.tare_baseline.cpu().numpy() if st is not None and st.tare_active else None
