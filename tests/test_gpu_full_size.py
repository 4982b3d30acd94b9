"""BASELINE.json configs 3, 4 and 5 at their FULL sizes on one B200, every output compared with the float64 oracle
(these sizes used to run only in builder-side scripts).  Config 2 at full size: test_gpu_kernel1.py.
Also the one real exchange of the design (config 4's row gather) on two ranks over NCCL when two GPUs are present."""
import os
import socket

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


def _welch_oracle(stream, n, hop, block=64):
    """O.welch_avg_peak_db with the per-segment transforms batched (linear_power_batch over blocks of segments):
    mean of |X_s|^2 over the segments (TraceAverager 'lin' with n >= nseg) and fmax of the per-segment dB."""
    w = O.make_window("hanning", n)
    nseg = (stream.shape[0] - n) // hop + 1
    total = np.zeros(n)
    peak = np.full(n, -np.inf)
    for s0 in range(0, nseg, block):
        s1 = min(s0 + block, nseg)
        seg = np.lib.stride_tricks.as_strided(stream[s0 * hop:], shape=(s1 - s0, n),
                                              strides=(hop * stream.itemsize, stream.itemsize))
        p = O.linear_power_batch(np.ascontiguousarray(seg), w, workers=-1)
        total += p.sum(axis=0)
        peak = np.maximum(peak, p.max(axis=0))
    return 10 * np.log10(total / nseg + O.POWER_LOG_FLOOR), 10 * np.log10(peak + O.POWER_LOG_FLOOR), nseg


CFG3_PATHS = {   # name -> environment switches of tdsa_welch (read at every call)
    "head_wl_tail": {},                                                   # default: radix-16 head kernel + warp-local tail kernel
    "fused": {"TDSA_WELCH_FUSED": "1"},                                   # both in ONE kernel, intermediate in an L2-resident ring
    "cluster": {"TDSA_WELCH_SUB": "0", "TDSA_WELCH_CLUSTER": "1"},        # round-1 16-CTA cluster kernel
    "two_kernel": {"TDSA_WELCH_SUB": "0", "TDSA_WELCH_CLUSTER": "0"},     # round-1 head + classic tail + linear rows
}
CFG3_KEYS = ("TDSA_WELCH_FUSED", "TDSA_WELCH_SUB", "TDSA_WELCH_CLUSTER")


def test_cfg3_full_size_welch_all_paths(dev, parity_log):
    """2^26 samples, N = 65536, hop 32768 -> 2047 segments: the default path (radix-16 head kernel, then 4096-point tails
    in fft_wl_kernel with the Welch state in tensor memory), the same in one fused launch, the 16-CTA cluster kernel and
    the round-1 two-kernel path against the oracle: the float64 plan at north_star's 1e-4 dB, the float32 plan at 1e-3 dB."""
    import torch
    from topdogspectrumanalyser_b200.engine import SpectrumPlan
    n, hop, total = 65536, 32768, 1 << 26
    stream = synth.cfg3_stream(n_samples=total, seed=2)
    want_avg, want_peak, nseg = _welch_oracle(stream, n, hop)
    assert nseg == 2047
    x = torch.from_numpy(stream).to(dev)
    keys = CFG3_KEYS
    old = {k: os.environ.get(k) for k in keys}
    try:
        for name, env in CFG3_PATHS.items():
            for k in keys:
                os.environ.pop(k, None)
            os.environ.update(env)
            for prec in ("f64", "f32"):
                plan = SpectrumPlan(n, precision=prec, device=dev)
                avg, peak = plan.welch(x, hop)
                avg2, peak2 = plan.welch(x, hop)              # a second call re-uses the re-armed claim counters
                assert torch.equal(avg, avg2) and torch.equal(peak, peak2)
                ea = float(np.abs(avg.cpu().numpy() - want_avg).max())
                ep = float(np.abs(peak.cpu().numpy() - want_peak).max())
                tol = TOL_DB if prec == "f64" else 1e-3       # float32 plan: measured 1.05e-4 dB on the peak row
                parity_log(f"cfg3_full_{name}_{prec}", max(ea, ep), tol=tol, avg_err=ea, peak_err=ep, segments=nseg)
                assert ea <= tol and ep <= tol, (name, prec, ea, ep)
                plan.close()
    finally:
        for k in keys:
            if old[k] is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = old[k]


@pytest.mark.parametrize("fused", ["1", "0"])
@pytest.mark.parametrize("nseg,hop", [(1, 16384), (3, 16384), (17, 16384), (40, 16384), (9, 12345)])
def test_welch_65536_few_segments(dev, nseg, hop, fused, monkeypatch):
    """Fewer segments than groups / CTAs with no work / odd counts: the fused and the two-launch path at 1e-4 dB.
    The odd hop puts segment starts on 8-byte boundaries only (the head kernel's 8-byte cp.async staging instead of its
    128-byte bulk copies)."""
    import torch
    from topdogspectrumanalyser_b200.engine import SpectrumPlan
    monkeypatch.setenv("TDSA_WELCH_FUSED", fused)
    n = 65536
    stream = synth.cfg3_stream(n_samples=n + hop * (nseg - 1) + 5, seed=11 + nseg)
    want_avg, want_peak, got_nseg = _welch_oracle(stream, n, hop)
    assert got_nseg == nseg
    plan = SpectrumPlan(n, device=dev)
    x = torch.from_numpy(stream).to(dev)
    for _ in range(2):                                    # the second call runs on re-armed counters
        avg, peak = plan.welch(x, hop)
        assert np.abs(avg.cpu().numpy() - want_avg).max() <= TOL_DB
        assert np.abs(peak.cpu().numpy() - want_peak).max() <= TOL_DB
    plan.close()


def test_welch_65536_more_segments_than_one_pass(dev, parity_log):
    """2053 heavily overlapping segments (hop 1024): the head + tail pair runs twice (2048 segments per pass of the
    intermediate buffer), the second pass with five segments on a smaller grid whose unused partial rows must read as
    empty; the fused variant takes all of them in one launch."""
    import torch
    from topdogspectrumanalyser_b200.engine import SpectrumPlan
    n, hop, nseg = 65536, 1024, 2053
    stream = synth.cfg3_stream(n_samples=n + hop * (nseg - 1), seed=91)
    want_avg, want_peak, got = _welch_oracle(stream, n, hop)
    assert got == nseg
    x = torch.from_numpy(stream).to(dev)
    old = os.environ.get("TDSA_WELCH_FUSED")
    try:
        for fused in ("0", "1"):
            os.environ["TDSA_WELCH_FUSED"] = fused
            plan = SpectrumPlan(n, device=dev)
            avg, peak = plan.welch(x, hop)
            ea = float(np.abs(avg.cpu().numpy() - want_avg).max()); ep = float(np.abs(peak.cpu().numpy() - want_peak).max())
            parity_log(f"welch65536_2053_segments_fused{fused}", max(ea, ep), tol=TOL_DB)
            assert ea <= TOL_DB and ep <= TOL_DB, (fused, ea, ep)
            plan.close()
    finally:
        if old is None:
            os.environ.pop("TDSA_WELCH_FUSED", None)
        else:
            os.environ["TDSA_WELCH_FUSED"] = old


@pytest.mark.parametrize("fused", ["0", "1"])
def test_welch_65536_psd_blackman(dev, fused, monkeypatch, parity_log):
    """PSD scaling (1 / (fs N)) and a different window through the config-3 path; second plan in float32 at 1e-3 dB."""
    import torch
    from topdogspectrumanalyser_b200.engine import SpectrumPlan
    monkeypatch.setenv("TDSA_WELCH_FUSED", fused)
    n, hop, fs = 65536, 20000, 20e6
    stream = synth.cfg3_stream(n_samples=n + hop * 24 + 17, seed=77)
    w = O.make_window("blackman", n)
    nseg = (stream.shape[0] - n) // hop + 1
    seg = np.lib.stride_tricks.as_strided(stream, shape=(nseg, n), strides=(hop * stream.itemsize, stream.itemsize))
    p = O.linear_power_batch(np.ascontiguousarray(seg), w, workers=-1) / (fs * n)
    want_avg = 10 * np.log10(p.mean(axis=0) + O.LOG_FLOOR)        # PSD mode: floor 1e-12 (rtl_samples.py:175-179)
    want_peak = 10 * np.log10(p.max(axis=0) + O.LOG_FLOOR)
    x = torch.from_numpy(stream).to(dev)
    for prec, tol in (("f64", TOL_DB), ("f32", 1e-3)):
        plan = SpectrumPlan(n, "blackman", mode="psd", fs=fs, precision=prec, device=dev)
        avg, peak = plan.welch(x, hop)
        ea = float(np.abs(avg.cpu().numpy() - want_avg).max()); ep = float(np.abs(peak.cpu().numpy() - want_peak).max())
        parity_log(f"welch65536_psd_blackman_fused{fused}_{prec}", max(ea, ep), tol=tol, segments=nseg)
        assert ea <= tol and ep <= tol, (prec, ea, ep)
        plan.close()


def test_cfg4_full_size_rows_and_grid(dev, parity_log):
    """300 sub-bands x 16 frames x 8192 points on one GPU: every dB row within 1e-4 dB, the 2 457 600-bin stitched grid
    bit-equal to the reference's argsort + np.interp (hackrf_sweep.py:150-166)."""
    import torch
    from topdogspectrumanalyser_b200.sweep import WidebandSweep
    nb, fr, n = 300, 16, 8192
    sw = WidebandSweep(nb, 20e6, n, 0.0, device=dev)
    assert sw.m == 2457600
    iq = synth.cfg4_subbands(nb, fr, n, seed=3)
    rows, grid = sw.run(torch.from_numpy(iq).to(dev))
    rows, grid = rows.cpu().numpy(), grid.cpu().numpy()
    w = O.make_window("hanning", n)
    worst = 0.0
    for b in range(nb):
        want = 10 * np.log10(O.linear_power_batch(iq[b], w).mean(axis=0) + O.POWER_LOG_FLOOR)
        worst = max(worst, float(np.abs(rows[b] - want).max()))
    los = [20e6 * i for i in range(nb)]
    want_grid = O.stitch_rows(rows, los, [lo + 20e6 for lo in los], O.sweep_grid(0, int(nb * 20e6), 20e6 / n))
    parity_log("cfg4_full_rows", worst, tol=TOL_DB, grid_bins=int(grid.size), grid_equal=bool(np.array_equal(grid, want_grid)))
    assert worst <= TOL_DB
    np.testing.assert_array_equal(grid, want_grid)
    # the sharded stitch produces the same elements, slice by slice
    from topdogspectrumanalyser_b200.engine import stitch
    from topdogspectrumanalyser_b200.sweep import shard_grid
    r = torch.from_numpy(rows).to(dev)
    for rank in (0, 3, 7):
        g0, cnt = shard_grid(sw.m, 8, rank)
        part = stitch(r, sw.band_lo_hz(range(nb)), sw.band_hz, sw.start_hz, sw.stop_hz, sw.m, g0, cnt).cpu().numpy()
        np.testing.assert_array_equal(part, want_grid[g0:g0 + cnt])


def test_cfg5_full_size_stream(dev, parity_log):
    """20 Msps for 10 s = 3052 chunks of 65 536 samples (48 832 frames of 4096), exp averaging n = 8, ring H = 1024:
    the ring's display view equals the oracle's ring after the same 48 832 frames."""
    from topdogspectrumanalyser_b200.streaming import WaterfallStreamer
    n, chunks, hist = 4096, 3052, 1024
    st = WaterfallStreamer(n_fft=n, chunk_samples=65536, history=hist, avg_mode="exp", avg_n=8, device=dev)
    stats = st.run(lambda c: synth.cfg5_chunk(c), chunks)
    assert stats["frames"] == chunks * 16
    a = O.TraceAverager()
    a.set_mode("exp", 8)
    w = O.make_window("hanning", n)
    ring = O.WaterfallRing(hist, n, -100.0)
    for c in range(chunks):
        p = O.linear_power_batch(synth.cfg5_chunk(c).reshape(16, n), w)
        for row in p:
            ring.add_row((10 * np.log10(a.process(row) + O.POWER_LOG_FLOOR)).astype(np.float32))
    err = float(np.abs(st.history().cpu().numpy() - ring.view()).max())
    parity_log("cfg5_full_ring_1024x4096", err, tol=TOL_DB, frames=chunks * 16, real_time_factor=stats["real_time_factor"])
    assert err <= TOL_DB


# ---- two ranks over NCCL ------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, exchange, q):
    import torch
    import torch.distributed as dist
    from topdogspectrumanalyser_b200.sweep import WidebandSweep, shard_bands, shard_grid
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        nb, fr, n = 21, 4, 8192                              # ragged shards: 11 + 10 sub-bands
        mine = shard_bands(nb, world, rank)
        iq = torch.from_numpy(synth.cfg4_subbands(nb, fr, n, seed=3, bands=mine)).to(dev)
        sw = WidebandSweep(nb, 20e6, n, 0.0, device=dev, exchange=exchange, grid="sharded")
        rows, part = sw.run(iq)
        rows2, part2 = sw.run(iq)                            # a second sweep through the same buffers
        torch.cuda.synchronize()
        g0, cnt = shard_grid(sw.m, world, rank)
        q.put((rank, sw.exchange, rows.cpu().numpy(), part.cpu().numpy(), g0, cnt,
               bool(torch.equal(rows, rows2) and torch.equal(part, part2))))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["nccl", "peer"])
def test_two_rank_row_exchange_and_sharded_stitch(exchange):
    """Config 4's one exchange on two GPUs: NCCL all-gather of the dB rows, and the fused variant in which the FFT
    kernel stores finished rows straight into both ranks' tables (peer memory); each rank stitches its grid slice."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, exchange, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=300) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    nb, fr, n = 21, 4, 8192
    iq = synth.cfg4_subbands(nb, fr, n, seed=3)
    w = O.make_window("hanning", n)
    want_rows = np.stack([10 * np.log10(O.linear_power_batch(b, w).mean(axis=0) + O.POWER_LOG_FLOOR) for b in iq])
    los = [20e6 * i for i in range(nb)]
    for rank, used, rows, part, g0, cnt, stable in results:
        assert used == exchange and stable
        assert np.abs(rows - want_rows).max() <= TOL_DB
        np.testing.assert_array_equal(rows, results[0][2])                       # both ranks hold the same table
        want_grid = O.stitch_rows(rows, los, [lo + 20e6 for lo in los], O.sweep_grid(0, int(nb * 20e6), 20e6 / n))
        np.testing.assert_array_equal(part, want_grid[g0:g0 + cnt])
