"""Round-2 kernels and paths: the accumulating (TMEM) epilogue of fft_wl_kernel, the two-engine N = 8192 kernel, the
device-resident flag block, the L2-chunked general scan, and the fixes for round-1 advisor findings.
Every check compares libtdsa.so (through the ctypes C ABI) with the float64 oracle on the same seeded IQ."""
import numpy as np
import pytest

from oracle import oracle as O
from topdogspectrumanalyser_b200 import synth

pytestmark = pytest.mark.gpu
TOL_DB = 1e-4          # north_star: every dB bin within 1e-4 of the float64 numpy chain


@pytest.fixture(scope="module")
def dev():
    import torch
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _oracle_avg_sequence(iq, n, mode, navg, averager=None):
    """Per-frame reference chain with TraceAverager (rtl_samples.py:169-184): rows, final buffer, max / min hold."""
    w = O.make_window("hanning", n)
    a = averager or O.TraceAverager()
    if averager is None:
        a.set_mode(mode, navg)
    rows, mx, mn = [], None, None
    for f in iq:
        db = O.power_db_frame(f, w, O.MODE_POWER, averager=a)
        rows.append(np.array(db, copy=True))
        mx = O.max_hold_update(mx, db.copy())
        mn = O.min_hold_update(mn, db.copy())
    return np.stack(rows), a, mx, mn


@pytest.mark.parametrize("mode,navg", [("exp", 8), ("lin", 100), ("lin", 1000)])
def test_fused_running_average_last_only(dev, parity_log, mode, navg):
    """Running average as a weighted sum in the FFT kernel's epilogue (N = 4096, last row only, no holds), across two
    calls so that the carry of a non-empty buffer is exercised; 'lin' both capped (n = 100 < frames) and uncapped."""
    import torch
    from topdogspectrumanalyser_b200.engine import SpectrumPlan, TraceState
    n, b1, b2 = 4096, 200, 150
    iq = synth.cfg2_frames(b=b1 + b2, n=n, seed=501)
    rows, a, _, _ = _oracle_avg_sequence(iq, n, mode, navg)
    plan = SpectrumPlan(n, device=dev)
    st = TraceState(n, dev)
    st.set_averaging(mode, navg)
    x = torch.from_numpy(iq).to(dev)
    r1 = plan.psd_db_avg_hold(x[:b1], st, last_only=True).cpu().numpy()
    e1 = np.abs(r1[0] - rows[b1 - 1]).max()
    r2 = plan.psd_db_avg_hold(x[b1:], st, last_only=True).cpu().numpy()
    e2 = np.abs(r2[0] - rows[-1]).max()
    buf = st.avg.cpu().numpy()
    rel = np.abs(buf - a._buffer).max() / np.abs(a._buffer).max()
    parity_log(f"fused_avg_{mode}{navg}_last_row", max(e1, e2), tol=TOL_DB, state_rel=float(rel))
    assert max(e1, e2) <= TOL_DB
    assert rel <= 1e-12                                   # the float64 state itself
    assert st.count == a._count and st.live_frames == b2
    plan.close()


@pytest.mark.parametrize("last_only", [False, True])
def test_fused_holds_without_averaging(dev, parity_log, last_only):
    """dB rows + max / min hold over a large un-averaged batch (N = 4096): hold = dB(max_f |X_f|^2), two calls."""
    import torch
    from topdogspectrumanalyser_b200.engine import SpectrumPlan, TraceState
    n, b1, b2 = 4096, 130, 90
    iq = synth.cfg2_frames(b=b1 + b2, n=n, seed=502)
    rows, _, mx, mn = _oracle_avg_sequence(iq, n, "off", 1)
    plan = SpectrumPlan(n, device=dev)
    st = TraceState(n, dev, max_hold_enabled=True, min_hold_enabled=True)
    x = torch.from_numpy(iq).to(dev)
    o1 = plan.psd_db_avg_hold(x[:b1], st, last_only=last_only).cpu().numpy()
    o2 = plan.psd_db_avg_hold(x[b1:], st, last_only=last_only).cpu().numpy()
    if last_only:
        err = max(np.abs(o1[0] - rows[b1 - 1]).max(), np.abs(o2[0] - rows[-1]).max())
    else:
        err = np.abs(np.concatenate([o1, o2]) - rows).max()
    eh = max(np.abs(st.max_hold.cpu().numpy() - mx).max(), np.abs(st.min_hold.cpu().numpy() - mn).max())
    parity_log(f"fused_holds_last_only={int(last_only)}", max(err, eh), tol=TOL_DB)
    assert err <= TOL_DB and eh <= TOL_DB
    assert st.valid == (1, 1)
    assert np.abs(st.last_row.cpu().numpy() - rows[-1]).max() <= TOL_DB
    plan.close()


@pytest.mark.parametrize("mode,navg", [("exp", 16), ("lin", 700), ("lin", 5000)])
def test_general_scan_in_l2_sized_chunks(dev, parity_log, mode, navg):
    """Averaging AND holds AND every row out: the frame-ordered scan (parallel over 32-frame blocks through the affine
    form of the recurrence), run over more frames than one 64 MB chunk of float64 rows holds (2048 frames at N = 4096),
    so the state is carried across block and chunk boundaries; 'lin' capped inside the run and never capped."""
    import torch
    from topdogspectrumanalyser_b200.engine import SpectrumPlan, TraceState
    n, b = 4096, 2300
    iq = synth.cfg2_frames(b=b, n=n, seed=503)
    rows, a, mx, mn = _oracle_avg_sequence(iq, n, mode, navg)
    plan = SpectrumPlan(n, device=dev)
    st = TraceState(n, dev, max_hold_enabled=True, min_hold_enabled=True)
    st.set_averaging(mode, navg)
    x = torch.from_numpy(iq).to(dev)
    got = torch.cat([plan.psd_db_avg_hold(x[:1500], st), plan.psd_db_avg_hold(x[1500:], st)]).cpu().numpy()
    err = np.abs(got - rows).max()
    eh = max(np.abs(st.max_hold.cpu().numpy() - mx).max(), np.abs(st.min_hold.cpu().numpy() - mn).max())
    rel = np.abs(st.avg.cpu().numpy() - a._buffer).max() / np.abs(a._buffer).max()
    parity_log(f"general_scan_chunked_{mode}{navg}_rows_holds", max(err, eh), tol=TOL_DB, frames=b, state_rel=float(rel))
    assert err <= TOL_DB and eh <= TOL_DB
    assert rel <= 1e-12
    assert st.count == a._count and st.live_frames == b - 1500
    assert np.abs(st.last_row.cpu().numpy() - rows[-1]).max() <= TOL_DB
    # last_only with a hold: same state, one row out
    st2 = TraceState(n, dev, max_hold_enabled=True)
    st2.set_averaging(mode, navg)
    last = plan.psd_db_avg_hold(x, st2, last_only=True).cpu().numpy()
    assert np.abs(last[0] - rows[-1]).max() <= TOL_DB
    assert np.abs(st2.max_hold.cpu().numpy() - mx).max() <= TOL_DB
    plan.close()


@pytest.mark.parametrize("prec,tol", [("f64", TOL_DB), ("f32", 5e-2)])
def test_two_engine_8192_kernel(dev, parity_log, prec, tol):
    """N = 8192 on the two-engine warp-local kernel (radix-2 DIF on the staged read, half-bin tables on the odd engine).
    It carries the group-mean epilogue; groups of ONE frame make it emit plain dB rows, compared bin by bin."""
    import torch
    from topdogspectrumanalyser_b200.engine import SpectrumPlan
    n, b = 8192, 333                                       # odd frame count, more than one wave of CTAs
    iq = synth.cfg2_frames(b=b, n=n, seed=504)
    x = torch.from_numpy(iq).to(dev)
    for window in ("hanning", "blackman"):
        want = O.power_db_batch(iq, O.make_window(window, n))
        plan = SpectrumPlan(n, window, precision=prec, device=dev)
        got = plan.group_avg_db(x.view(b, 1, n)).cpu().numpy().astype(np.float64)
        err = np.abs(got - want)
        parity_log(f"wl8192_{prec}_{window}", err.max(), tol=tol, over_1e4=int((err > 1e-4).sum()), bins=int(err.size))
        assert err.max() <= tol
        # plain dB rows (float64: the same two-engine kernel with a row epilogue; float32: the four-pass classic kernel)
        classic = plan.psd_db(x).cpu().numpy().astype(np.float64)
        assert np.abs(classic - want).max() <= tol
        assert np.abs(classic - got).max() <= (2e-5 if prec == "f64" else tol)
        lin = plan.power_linear(x[:40]).cpu().numpy()            # float64 linear rows (the general trace path's input)
        assert np.abs(10 * np.log10(lin + O.POWER_LOG_FLOOR) - want[:40]).max() <= tol
        plan.close()
    plan = SpectrumPlan(n, "hanning", mode="psd", fs=20e6, precision=prec, device=dev)
    want = O.power_db_batch(iq[:64], O.make_window("hanning", n), O.MODE_PSD, fs=20e6)
    got = plan.group_avg_db(x[:64].view(64, 1, n)).cpu().numpy()
    assert np.abs(got - want).max() <= tol
    plan.close()
    # overlapping frames of a flat stream (stride 1000 samples) in mag20 mode
    stream = synth.cfg3_stream(n_samples=8192 + 1000 * 36, seed=507)
    frames = np.lib.stride_tricks.as_strided(stream, shape=(37, n), strides=(1000 * stream.itemsize, stream.itemsize))
    want = O.power_db_batch(np.ascontiguousarray(frames), O.make_window("hanning", n), O.MODE_MAG20)
    plan = SpectrumPlan(n, "hanning", mode="mag20", precision=prec, device=dev)
    got = plan.psd_db(torch.from_numpy(stream).to(dev), n_frames=37, frame_stride=1000).cpu().numpy()
    assert np.abs(got - want).max() <= tol
    # an odd stride cannot be described by the tensor map: the four-pass classic kernel serves these rows
    frames = np.lib.stride_tricks.as_strided(stream, shape=(36, n), strides=(1001 * stream.itemsize, stream.itemsize))
    want = O.power_db_batch(np.ascontiguousarray(frames), O.make_window("hanning", n), O.MODE_MAG20)
    got = plan.psd_db(torch.from_numpy(stream).to(dev), n_frames=36, frame_stride=1001).cpu().numpy()
    assert np.abs(got - want).max() <= tol
    plan.close()


@pytest.mark.parametrize("n,groups,frames", [(4096, 37, 16), (8192, 21, 16), (8192, 5, 3), (8192, 300, 16), (4096, 700, 4)])
def test_group_mean_in_the_epilogue(dev, parity_log, n, groups, frames):
    """Config-4 rows: the linear mean of each group of frames, formed in TMEM inside the FFT kernel, as one dB row."""
    import torch
    from topdogspectrumanalyser_b200.engine import SpectrumPlan
    iq = synth.cfg4_subbands(n_bands=groups, frames=frames, n=n, seed=505)
    w = O.make_window("hanning", n)
    want = []
    for band in iq:
        a = O.TraceAverager()
        a.set_mode("lin", frames)
        for f in band:
            db = O.power_db_frame(f, w, O.MODE_POWER, averager=a)
        want.append(np.array(db, copy=True))
    plan = SpectrumPlan(n, device=dev)
    got = plan.group_avg_db(torch.from_numpy(iq).to(dev)).cpu().numpy()
    err = np.abs(got - np.stack(want)).max()
    parity_log(f"group_mean_n{n}_g{groups}x{frames}", err, tol=TOL_DB)
    assert err <= TOL_DB
    plan.close()


def test_welch_4096_in_the_epilogue(dev, parity_log):
    """Welch mean + peak with overlapping 4096-point segments: sum and max accumulate in TMEM (one launch)."""
    import torch
    from topdogspectrumanalyser_b200.engine import SpectrumPlan
    n, hop, total = 4096, 2048, 1 << 21                    # 1023 segments
    stream = synth.cfg3_stream(n_samples=total, seed=506)
    want_avg, want_peak = O.welch_avg_peak_db(stream, O.make_window("hanning", n), hop)
    plan = SpectrumPlan(n, device=dev)
    avg, peak = plan.welch(torch.from_numpy(stream).to(dev), hop)
    err = max(np.abs(avg.cpu().numpy() - want_avg).max(), np.abs(peak.cpu().numpy() - want_peak).max())
    parity_log("welch_4096_fused", err, tol=TOL_DB, segments=(total - n) // hop + 1)
    assert err <= TOL_DB
    plan.close()


@pytest.mark.parametrize("n", [16384, 65536])
def test_hackrf_front_end_on_large_transforms(dev, parity_log, n):
    """the DC estimates live in their own allocation, so the large-FFT path (which keeps its
    scratch in scratch2) no longer overwrites them. psd_db_dc and avg_hold_dc at N >= 16384 with several frames."""
    import torch
    from topdogspectrumanalyser_b200.engine import SpectrumPlan, TraceState
    b, fs = 5, 20e6
    rng = np.random.default_rng(507 + n)
    iq = synth.cfg2_frames(b=b, n=n, seed=508)
    iq = (iq + (0.3 - 0.2j) + rng.standard_normal((b, 1)).astype(np.float32) * 0.05).astype(np.complex64)   # a DC offset per frame
    iq[2] = 0                                                                                                 # one silent frame
    win = O.make_window_hackrf(n)

    def truth(averager):
        dc, rows, last = 0.0 + 0.0j, [], None
        for f in iq:
            db, dc = O.hackrf_power_db_frame(f, win, use_psd=False, fs=fs, averager=averager, dc_estimate=dc)
            last = last if db is None else np.array(db, copy=True)
            rows.append(last)
        return np.stack(rows), dc

    want, dc_final = truth(None)
    plan = SpectrumPlan(n, "hanning", "rms", "mag20", fs=fs, device=dev)
    dcs = torch.zeros(2, dtype=torch.float64, device=dev)
    x = torch.from_numpy(iq).to(dev)
    got, silent = plan.psd_db_dc(x, dcs)
    got = got.cpu().numpy()
    assert silent.cpu().tolist() == [0, 0, 1, 0, 0]
    live = [0, 1, 3, 4]
    worst = float(np.abs(got[live] - want[live]).max())
    parity_log(f"hackrf_dc_front_end_mag20_n{n}", worst, tol=TOL_DB)
    assert worst <= TOL_DB
    d = dcs.cpu().numpy()
    assert abs(d[0] - dc_final.real) <= 1e-12 and abs(d[1] - dc_final.imag) <= 1e-12
    # averaging variant through the same front end; the silent frame repeats the previous row and is not counted
    plan.set_mode("power")
    st = TraceState(n, dev)
    st.set_averaging("exp", 4)
    dcs.zero_()
    rows, silent = plan.psd_db_avg_hold_dc(x, st, dcs)
    rows = rows.cpu().numpy()
    a = O.TraceAverager()
    a.set_mode("exp", 4)
    want, _ = truth(a)
    worst = float(np.abs(rows - want).max())
    parity_log(f"hackrf_dc_avg_exp4_n{n}", worst, tol=TOL_DB)
    assert worst <= TOL_DB and st.live_frames == 4
    assert np.array_equal(rows[2], rows[1])
    plan.close()


def test_silent_frame_at_the_start_of_a_call_repeats_the_last_good_row(dev):
    """the last good dB row is carried across calls (hackrf_samples.py:351-355)."""
    import torch
    from topdogspectrumanalyser_b200.engine import SpectrumPlan, TraceState
    n = 1024
    iq = synth.cfg2_frames(b=3, n=n, seed=509)
    plan = SpectrumPlan(n, "hanning", "rms", "power", device=dev)
    st = TraceState(n, dev)
    st.set_averaging("exp", 4)
    dcs = torch.zeros(2, dtype=torch.float64, device=dev)
    first, _ = plan.psd_db_avg_hold_dc(torch.from_numpy(iq[:2]).to(dev), st, dcs)
    quiet = np.zeros((2, n), dtype=np.complex64)
    quiet[1] = iq[2]
    second, silent = plan.psd_db_avg_hold_dc(torch.from_numpy(quiet).to(dev), st, dcs)
    assert silent.cpu().tolist() == [1, 0]
    np.testing.assert_array_equal(second[0].cpu().numpy(), first[1].cpu().numpy())     # not a row of zeros
    only, _ = plan.psd_db_avg_hold_dc(torch.from_numpy(quiet[:1]).to(dev), st, dcs, last_only=True)
    np.testing.assert_array_equal(only[0].cpu().numpy(), second[1].cpu().numpy())
    assert st.live_frames == 0 and st.count == 1
    plan.close()


def test_host_scalar_entry_points_still_agree(dev):
    """tdsa_psd_db_avg_hold / tdsa_trace_update keep their host-scalar contract on top of the device flag block."""
    import ctypes as C
    import torch
    from topdogspectrumanalyser_b200 import _lib as L
    from topdogspectrumanalyser_b200.engine import SpectrumPlan
    lib = L.load()
    n, b = 1024, 12
    iq = synth.cfg2_frames(b=b, n=n, seed=510)
    rows, a, mx, mn = _oracle_avg_sequence(iq, n, "lin", 5)
    plan = SpectrumPlan(n, device=dev)
    plan._bind()
    x = torch.from_numpy(iq).to(dev)
    avg = torch.zeros(n, dtype=torch.float64, device=dev)
    hmx, hmn = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    out = torch.empty((b, n), device=dev)
    count, valid = C.c_int32(0), (C.c_int32 * 2)(0, 0)
    for lo, hi in ((0, 7), (7, b)):
        L.check(lib.tdsa_psd_db_avg_hold(plan._h, x[lo:hi].data_ptr(), hi - lo, n, L.AVG_LIN, 5, avg.data_ptr(), C.byref(count),
                                         hmx.data_ptr(), hmn.data_ptr(), valid, 0, out[lo:hi].data_ptr()))
    assert count.value == a._count == 5 and list(valid) == [1, 1]
    assert np.abs(out.cpu().numpy() - rows).max() <= TOL_DB
    assert np.abs(hmx.cpu().numpy() - mx).max() <= TOL_DB and np.abs(hmn.cpu().numpy() - mn).max() <= TOL_DB
    plan.close()


def test_waterfall_ring_dedupe_and_image(dev, golden):
    """The widget's duplicate filter (displays/waterfall.py:330-336) and the exporter's RGBA image
    (core/export_manager.py:72-79) against the oracle ring driven the way update_widget_data drives it."""
    import torch
    from topdogspectrumanalyser_b200.engine import WaterfallRing
    g = golden("waterfall_ring.npz")
    h, w = int(g["h"]), g["rows"].shape[1]
    rows = g["rows"].copy()
    # the 20 ms tick re-delivers rows: repeat some, put a NaN row twice (NaN never compares equal: both are added)
    seq = [0, 0, 1, 2, 2, 2, 3, 4, 4, 5, 6, 6, 7, 8, 9, 9, 10, 11, 12, 13, 13]
    feed = rows[seq].copy()
    feed[7, 3] = np.nan
    feed[8] = feed[7]
    ring = WaterfallRing(h, w, float(g["fill"]), dev, dedupe=True)
    want = O.WaterfallRing(h, w, float(g["fill"]))
    last = None
    x = torch.from_numpy(feed).to(dev)
    i = 0
    for step in (1, 3, 2, 7, 1, 7):
        added = 0
        for r in feed[i:i + step]:
            if last is None or not np.array_equal(r, last):
                last = r.copy()
                want.add_row(r)
                added += 1
        ring.push(x[i:i + step])
        i += step
        np.testing.assert_array_equal(ring.view().cpu().numpy(), want.view())
        assert ring.rows_added == added
    # the colour-mapped image of the display view
    lut = (np.arange(256)[:, None] * np.array([1, 2, 3, 0]) % 256 + np.array([0, 0, 0, 255])).astype(np.uint8)
    lo, hi = -90.0, -20.0
    img = ring.image(lo, hi, torch.from_numpy(lut).to(dev)).cpu().numpy()
    arr = want.view().astype(np.float32)
    with np.errstate(invalid="ignore"):
        norm = np.clip((arr - np.float32(lo)) / np.float32(max(hi - lo, 1e-9)), 0.0, 1.0)
        idx = np.nan_to_num(norm * 255, nan=0.0).astype(np.uint8)
    np.testing.assert_array_equal(img, lut[idx])
    # a plain (host pointer) ring can hand out an image too
    plain = WaterfallRing(h, w, float(g["fill"]), dev)
    plain.push(torch.from_numpy(rows[:5]).to(dev))
    ref = O.WaterfallRing(h, w, float(g["fill"]))
    for r in rows[:5]:
        ref.add_row(r)
    img2 = plain.image(lo, hi, torch.from_numpy(lut).to(dev)).cpu().numpy()
    norm = np.clip((ref.view().astype(np.float32) - np.float32(lo)) / np.float32(hi - lo), 0.0, 1.0)
    np.testing.assert_array_equal(img2, lut[(norm * 255).astype(np.uint8)])


def test_snap_to_peak_matches_scipy_find_peaks(dev):
    """core/marker_manager.py:74-99 against scipy.signal.find_peaks itself."""
    import torch
    from scipy.signal import find_peaks
    from topdogspectrumanalyser_b200.analytics import snap_to_peak
    rng = np.random.default_rng(511)
    for trial in range(12):
        n = int(rng.choice([512, 1024, 4096, 16384]))
        x = (rng.standard_normal(n) * 3.0 - 90.0).astype(np.float32)
        for _ in range(int(rng.integers(0, 6))):                    # a few carriers, some closer than `distance`
            k = int(rng.integers(2, n - 3))
            x[k] += float(rng.uniform(5, 40))
            if rng.random() < 0.5:
                x[k + 2] = x[k] - np.float32(0.5)
        if trial % 4 == 1:                                          # a flat-topped peak: the plateau's midpoint counts
            k = int(rng.integers(10, n - 20))
            x[k:k + 5] = x.max() + np.float32(7.0)
        if trial % 4 == 2:
            x[:] = np.sort(x)                                       # monotonic: no peak at all -> argmax fallback
        thr, exc = (-200.0, 6.0) if trial % 3 else (-85.0, 12.0)
        peaks, props = find_peaks(x, height=thr, prominence=exc, distance=3)
        want = int(peaks[int(np.argmax(props["peak_heights"]))]) if len(peaks) else int(np.argmax(x))
        bins = np.linspace(88e6, 108e6, n)
        f, idx, fb = snap_to_peak(bins, torch.from_numpy(x).to(dev), thr, exc, 3)
        assert idx == want and f == bins[want] and fb == (len(peaks) == 0), (trial, n, idx, want)


def test_iq_in_sweep_source(dev, parity_log):
    """B200SweepDataSource: the job of the external hackrf_sweep binary done from raw IQ. Rows = kernel 1 per frame +
    linear mean (float64 oracle, 1e-4 dB); the stitched grid is what HackRFSweepDataSource._parse builds from those rows
    (argsort + np.interp onto linspace(start, stop, int(span / bin_size))), bit for bit."""
    from topdogspectrumanalyser_b200.datasources import B200SweepDataSource, SweepDataSource, SyntheticTunerFeed
    start, stop, bin_size = 2400e6, 2500e6, 30000
    feed = SyntheticTunerFeed(20e6, carriers_hz=[2.412e9, 2.437e9, 2.4835e9], amp=0.7, seed=21)
    src = B200SweepDataSource(start, stop, bin_size, feed=feed, frames=4)
    assert isinstance(src, SweepDataSource) and src.n_fft == 1024 and src.n_bands == 5
    assert np.isnan(src.get_data()).all() and src.get_number_of_points() == 3333       # nothing swept yet
    iq = src.acquire_sweep()
    grid = src.process_sweep(iq)
    w = O.make_window("hanning", src.n_fft)
    rows = np.stack([10 * np.log10(O.linear_power_batch(b, w).mean(axis=0) + O.POWER_LOG_FLOOR) for b in iq])
    got_rows = src._plan.group_avg_db(src._dev_iq).cpu().numpy()
    err = float(np.abs(got_rows - rows).max())
    parity_log("iq_sweep_rows", err, tol=TOL_DB)
    assert err <= TOL_DB
    los = [start + 20e6 * i for i in range(5)]
    want = O.stitch_rows(got_rows, los, [lo + 20e6 for lo in los], O.sweep_grid(int(start), int(stop), bin_size))
    np.testing.assert_array_equal(grid, want)
    np.testing.assert_array_equal(src.get_data(), want)
    assert src.get_data() is not src.get_data()                                         # a copy per call, like the reference
    # the three carriers stand out where they should
    for f in (2.412e9, 2.437e9, 2.4835e9):
        k = int(np.argmin(np.abs(src.frequency_grid - f)))
        assert grid[max(k - 3, 0):k + 4].max() > np.nanmedian(grid) + 20
    # background thread: start / a few sweeps / stop
    src.start()
    import time
    t0 = time.time()
    while src.sweep_rate is None and time.time() - t0 < 20:
        time.sleep(0.05)
    src.stop()
    assert src.sweep_rate is not None and not src.is_running and not np.isnan(src.get_data()).any()
