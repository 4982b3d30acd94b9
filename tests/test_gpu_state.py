"""Trace state, data source, stitch, ring, Welch and streaming: device results vs oracle / golden fixtures."""
import numpy as np
import pytest

from oracle import oracle as O
from topdogspectrumanalyser_b200 import synth

pytestmark = pytest.mark.gpu
TOL_DB = 1e-4


@pytest.fixture(scope="module")
def dev():
    import torch
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


from topdogspectrumanalyser_b200.datasources import ReplayFeed  # noqa: E402


def make_source(frames, n, fs, fc, **kw):
    from topdogspectrumanalyser_b200.datasources import B200SampleDataSource
    src = B200SampleDataSource(int(fs), int(fc), feed=ReplayFeed(frames, fs, fc), **kw)
    src.set_fft_size(n)
    src.start()
    return src


def test_datasource_cfg1_matches_reference_object(dev, golden):
    """BASELINE config 1 through the boundary class: same arrays as RtlSamplesDataSource.get_power_levels."""
    g = golden("rtl_chain.npz")
    fs, fc = float(g["fs"]), float(g["fc"])
    src = make_source(g["cfg1_iq"], 1024, fs, fc)
    outs = []
    for want in g["cfg1_db"]:
        p, bins = src.get_power_levels()
        assert p.dtype == np.float64 and p.shape == (1024,)
        np.testing.assert_array_equal(bins, g["cfg1_bins"])
        assert np.abs(p - want).max() <= TOL_DB
        outs.append(p)
    assert all(a is not b for a, b in zip(outs, outs[1:]))       # fresh array every call (aliasing contract)
    assert src.get_raw_samples() is not None and src.last_data_time > 0
    assert src.sample_count == 1024 and src.num_samples == 1024 and src.window_type == "hanning"


@pytest.mark.parametrize("key,psd,avg", [("avg_exp8_power", False, ("exp", 8)), ("avg_lin4_power", False, ("lin", 4)),
                                         ("avg_lin4_psd", True, ("lin", 4))])
def test_datasource_averaging_sequences(dev, golden, key, psd, avg):
    g = golden("rtl_chain.npz")
    src = make_source(g["avg_iq"], 512, float(g["fs"]), float(g["fc"]))
    src.set_psd_mode(psd)
    src.set_averaging(*avg)
    for want in g[key]:
        p, _ = src.get_power_levels()
        assert np.abs(p - want).max() <= TOL_DB
    # reset_averaging restarts the running mean: next frame equals the un-averaged spectrum
    src.sdr.i = 0
    src.reset_averaging()
    p, _ = src.get_power_levels()
    assert np.abs(p - g[key][0]).max() <= TOL_DB


def test_datasource_setters_follow_reference_semantics(dev, golden):
    g = golden("rtl_chain.npz")
    fs, fc = float(g["fs"]), float(g["fc"])
    src = make_source(g["w_iq"], 4096, fs, fc)
    src.set_window_type("hamming")
    p, _ = src.get_power_levels()
    assert np.abs(p - g["w_hamming_power"][0]).max() <= TOL_DB
    src.set_psd_mode(True)
    p, _ = src.get_power_levels()
    assert np.abs(p - g["w_hamming_psd"][1]).max() <= TOL_DB
    src.set_window_type("no-such-window")                        # falls back to hanning (rtl_samples.py:205)
    assert src.window_type == "hanning"
    src.sample_count = 2048                                      # size change: Hann again, bins follow
    assert src.fft_size == 2048 and src.window_type == "hanning"
    src.sdr = ReplayFeed(g["n2048_iq"], fs, fc)
    src.set_psd_mode(False)
    p, bins = src.get_power_levels()
    assert len(p) == len(bins) == 2048
    assert np.abs(p - g["n2048_db"][0]).max() <= TOL_DB
    # not running -> zeros + linspace bins, never raises (rtl_samples.py:149-155)
    src.pause()
    p, bins = src.get_power_levels()
    assert not p.any() and len(bins) == 2048
    # a feed that throws -> logged, zeros returned (rtl_samples.py:191-197)
    src.resume()
    src.sdr = ReplayFeed([], fs, fc)
    p, bins = src.get_power_levels()
    assert not p.any()


def test_datasource_hackrf_style(dev, golden, parity_log):
    """DC removal + RMS-normalised float32 Hann + the three dB branches + silence hold."""
    g = golden("hackrf_chain.npz")
    fs, fc = float(g["fs"]), float(g["fc"])
    win = O.make_window_hackrf(1024)

    def truth(use_psd, avg):
        a = O.TraceAverager()
        if avg:
            a.set_mode(*avg)
        last, rows = None, []
        for iq in g["iq"]:
            db, _ = O.hackrf_power_db_frame(iq, win, use_psd=use_psd, fs=fs, averager=a)
            db = last if db is None else db
            last = db
            rows.append(np.array(db, copy=True))
        return np.stack(rows)

    for use_psd, avg, key in ((False, None, "mag20"), (True, None, "psd"), (False, ("exp", 4), "avg_exp4")):
        src = make_source(g["iq"], 1024, fs, fc, style="hackrf")
        src.sdr.dtype = np.complex64
        src.set_psd_mode(use_psd)
        if avg:
            src.set_averaging(*avg)
        want = truth(use_psd, avg)
        worst = 0.0
        for i in range(len(want)):
            p, bins = src.get_power_levels()
            assert p.dtype == np.float32
            worst = max(worst, float(np.abs(p.astype(np.float64) - want[i]).max()))
            strong = g[key][i] > g[key][i].max() - 60
            assert np.abs(p - g[key][i])[strong].max() < 5e-3                       # executed reference (f32 FFT)
        parity_log(f"hackrf_chain_{key}_vs_f64_oracle", worst, tol=TOL_DB)
        assert worst <= TOL_DB, (key, worst)                                        # north_star's bar, no slack
        np.testing.assert_array_equal(bins, g["bins"])


def test_avg_hold_batch_matches_sequential_reference(dev):
    import torch
    from topdogspectrumanalyser_b200.engine import SpectrumPlan, TraceState
    n, b = 1024, 24
    iq = synth.cfg2_frames(b=b, n=n, seed=77)
    w = O.make_window("hanning", n)
    for mode, navg in (("exp", 8), ("lin", 5), ("off", 1)):
        a = O.TraceAverager()
        a.set_mode(mode, navg)
        want, mx, mn = [], None, None
        for f in iq:
            db = O.power_db_frame(f, w, O.MODE_POWER, averager=a)
            want.append(np.array(db, copy=True))
            mx = O.max_hold_update(mx, db.copy())
            mn = O.min_hold_update(mn, db.copy())
        want = np.stack(want)
        plan = SpectrumPlan(n, device=dev)
        st = TraceState(n, dev)
        st.set_averaging(mode, navg)
        st.max_hold_enabled = st.min_hold_enabled = True
        x = torch.from_numpy(iq).to(dev)
        got = torch.cat([plan.psd_db_avg_hold(x[:10], st), plan.psd_db_avg_hold(x[10:], st)]).cpu().numpy()
        assert np.abs(got - want).max() <= TOL_DB, mode
        assert np.abs(st.max_hold.cpu().numpy() - mx).max() <= TOL_DB
        assert np.abs(st.min_hold.cpu().numpy() - mn).max() <= TOL_DB
        last = plan.psd_db_avg_hold(x[:1], st, last_only=True)
        assert last.shape == (1, n)
        plan.close()


def test_trace_update_holds_nan_and_sweep_average(dev, golden):
    import torch
    from topdogspectrumanalyser_b200.engine import TraceState, trace_update
    g = golden("trace_state.npz")
    frames = g["frames"][:8].astype(np.float32)
    st = TraceState(frames.shape[1], dev)
    st.max_hold_enabled = st.min_hold_enabled = True
    for i in range(0, 8, 3):                                    # state carries across calls
        trace_update(torch.from_numpy(frames[i:i + 3]).to(dev), st)
        hi = min(i + 3, 8) - 1
        np.testing.assert_allclose(st.max_hold.cpu().numpy(), g["max_hold"][hi].astype(np.float32), rtol=0, atol=1e-5)
        np.testing.assert_allclose(st.min_hold.cpu().numpy(), g["min_hold"][hi].astype(np.float32), rtol=0, atol=1e-5)
    # sweep-domain averaging exp n=4 on dB rows, with an all-NaN frame that must be skipped
    rows = g["tare_in"][:10].astype(np.float32)
    rows_nan = np.insert(rows, 4, np.nan, axis=0)
    st = TraceState(rows.shape[1], dev)
    st.set_averaging("exp", 4)
    out = trace_update(torch.from_numpy(np.ascontiguousarray(rows_nan)).to(dev), st).cpu().numpy()
    got = np.delete(out, 4, axis=0)
    a = O.TraceAverager()
    a.set_mode("exp", 4)
    want = np.stack([O.sweep_average_db(r.astype(np.float64), a) for r in rows])
    assert np.abs(got - want).max() <= 2e-5
    assert np.isnan(out[4]).all()
    # calibration offset is a plain dB add (display_data_processor.py:317-327)
    st = TraceState(rows.shape[1], dev)
    out = trace_update(torch.from_numpy(rows).to(dev), st, cal_offset_db=-3.25).cpu().numpy()
    np.testing.assert_allclose(out, rows - 3.25, rtol=0, atol=1e-5)


def test_stitch_matches_reference_parse(dev, golden):
    import torch
    from topdogspectrumanalyser_b200.engine import stitch
    g = golden("sweep_stitch.npz")
    rows = torch.from_numpy(g["rows"]).to(dev)                  # arrival order (interleaved), like hackrf_sweep
    lo = torch.from_numpy(g["lo"].astype(np.float64)).to(dev)
    got = stitch(rows, lo, float(g["hi"][0] - g["lo"][0]), float(g["start"]), float(g["stop"]), len(g["grid"]))
    np.testing.assert_array_equal(got.cpu().numpy(), g["stitched"])     # float64, bit-exact


def test_waterfall_ring(dev, golden):
    import torch
    from topdogspectrumanalyser_b200.engine import WaterfallRing
    g = golden("waterfall_ring.npz")
    h, w = int(g["h"]), g["rows"].shape[1]
    ring = WaterfallRing(h, w, float(g["fill"]), dev)
    rows = torch.from_numpy(g["rows"]).to(dev)
    i = 0
    for step in (1, 2, 5, 1, 13, 8):                            # pushes of several rows at once, one > H
        ring.push(rows[i:i + step])
        i += step
        np.testing.assert_array_equal(ring.view().cpu().numpy(), g["views"][i - 1])


@pytest.mark.parametrize("n,hop,total", [(4096, 2048, 1 << 17), (65536, 32768, 1 << 19)])
def test_welch_avg_and_peak(dev, n, hop, total):
    import torch
    from topdogspectrumanalyser_b200.engine import SpectrumPlan
    stream = synth.cfg3_stream(n_samples=total, seed=2)
    want_avg, want_peak = O.welch_avg_peak_db(stream, O.make_window("hanning", n), hop)
    plan = SpectrumPlan(n, device=dev)
    avg, peak = plan.welch(torch.from_numpy(stream).to(dev), hop)
    assert np.abs(avg.cpu().numpy() - want_avg).max() <= TOL_DB
    assert np.abs(peak.cpu().numpy() - want_peak).max() <= TOL_DB
    plan.close()


def test_group_avg_and_wideband_stitch_single_gpu(dev):
    """Config 4 at reduced size on one GPU: 6 sub-bands x 4 frames x 8192, stitched like the reference."""
    import torch
    from topdogspectrumanalyser_b200.sweep import WidebandSweep
    nb, frames, n = 6, 4, 8192
    iq = synth.cfg4_subbands(n_bands=nb, frames=frames, n=n, seed=3)
    sw = WidebandSweep(n_bands=nb, band_hz=20e6, n_fft=n, start_hz=0.0, device=dev)
    rows, grid = sw.run(torch.from_numpy(iq).to(dev))
    w = O.make_window("hanning", n)
    want_rows = []
    for band in iq:
        a = O.TraceAverager()
        a.set_mode("lin", frames)
        for f in band:
            db = O.power_db_frame(f, w, O.MODE_POWER, averager=a)
        want_rows.append(np.array(db, copy=True))
    want_rows = np.stack(want_rows)
    got_rows = rows.cpu().numpy()
    assert np.abs(got_rows - want_rows).max() <= TOL_DB
    los = [20e6 * i for i in range(nb)]
    his = [lo + 20e6 for lo in los]
    want_grid = O.stitch_rows(got_rows, los, his, O.sweep_grid(0, int(nb * 20e6), 20e6 / n))
    assert sw.m == len(want_grid)
    np.testing.assert_array_equal(grid.cpu().numpy(), want_grid)


def test_streaming_waterfall(dev):
    """Config 5 shape at reduced length: chunks of 65 536 samples, exp avg n=8, ring of the last H rows."""
    from topdogspectrumanalyser_b200.streaming import WaterfallStreamer
    n, chunks, hist = 4096, 6, 64
    st = WaterfallStreamer(n_fft=n, chunk_samples=65536, history=hist, avg_mode="exp", avg_n=8, device=dev)
    stats = st.run(lambda c: synth.cfg5_chunk(c), chunks)
    assert stats["frames"] == chunks * 16
    a = O.TraceAverager()
    a.set_mode("exp", 8)
    w = O.make_window("hanning", n)
    ring = O.WaterfallRing(hist, n, -100.0)
    for c in range(chunks):
        for f in synth.cfg5_chunk(c).reshape(16, n):
            ring.add_row(O.power_db_frame(f, w, O.MODE_POWER, averager=a).astype(np.float32))
    assert np.abs(st.history().cpu().numpy() - ring.view()).max() <= TOL_DB


def test_cabi_error_codes_and_messages(dev):
    """Every entry point answers bad arguments with a negative TDSA_ERR_* and a message (include/tdsa.h conventions)."""
    import ctypes as C
    import torch
    from topdogspectrumanalyser_b200 import _lib as L
    from topdogspectrumanalyser_b200.engine import SpectrumPlan, TraceState
    lib = L.load()
    h = C.c_void_p()
    assert lib.tdsa_create(1000, 0, 0, 0, 1e-10, 1.0, 0, C.byref(h)) == -2 and b"powers of two" in lib.tdsa_last_error()
    assert lib.tdsa_create(1024, 9, 0, 0, 1e-10, 1.0, 0, C.byref(h)) == -1
    assert lib.tdsa_create(1024, 0, 0, 7, 1e-10, 1.0, 0, C.byref(h)) == -1
    assert lib.tdsa_create(1024, 0, 0, 0, 1e-10, 1.0, 5, C.byref(h)) == -1
    assert lib.tdsa_create(1 << 21, 0, 0, 0, 1e-10, 1.0, 0, C.byref(h)) == -2
    assert lib.tdsa_psd_db_batch(None, None, 1, 1024, None) == -1
    plan = SpectrumPlan(1024, device=dev)
    x = torch.zeros((2, 1024), dtype=torch.complex64, device=dev)
    y = torch.zeros((2, 1024), dtype=torch.float32, device=dev)
    assert lib.tdsa_psd_db_batch(plan._h, x.data_ptr(), -1, 1024, y.data_ptr()) == -1
    assert lib.tdsa_psd_db_batch(plan._h, x.data_ptr(), 2, 0, y.data_ptr()) == -1
    assert lib.tdsa_psd_db_batch(plan._h, None, 2, 1024, y.data_ptr()) == -1
    assert lib.tdsa_psd_db_batch(plan._h, x.data_ptr() + 4, 1, 1024, y.data_ptr()) == -1 and b"aligned" in lib.tdsa_last_error()
    assert lib.tdsa_psd_db_batch(plan._h, x.data_ptr(), 0, 1024, y.data_ptr()) == 0          # empty batch is fine
    assert lib.tdsa_welch(plan._h, x.data_ptr(), 100, 512, y.data_ptr(), y.data_ptr()) == -1
    assert lib.tdsa_set_mode(plan._h, 9, 1e-10, 1.0) == -1
    # the mag20 branch is never averaged in the reference (hackrf_samples.py:378-383)
    plan.set_mode("mag20")
    st = TraceState(1024, dev)
    st.set_averaging("exp", 4)
    with pytest.raises(L.TdsaError):
        plan.psd_db_avg_hold(x, st)
    # averaging without state / holds without flags
    plan.set_mode("power")
    assert lib.tdsa_psd_db_avg_hold(plan._h, x.data_ptr(), 2, 1024, 1, 4, None, None, None, None, None, 0, y.data_ptr()) == -1
    assert lib.tdsa_top_peaks(y.data_ptr(), 1 << 20, 5, 10, C.c_float(10.0), None, None, None, None) == -1
    idx = torch.zeros(16, dtype=torch.int32, device=dev)
    assert lib.tdsa_top_peaks(y.data_ptr(), 1 << 20, 5, 10, C.c_float(10.0), idx.data_ptr(), y.data_ptr(), idx.data_ptr(), None) == -2
    plan.close()
    plan.close()                                                  # idempotent
