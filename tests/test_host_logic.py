"""Host-side logic that needs no GPU: sharding, gloo all-gather (world_size 2), synthetic generators."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from topdogspectrumanalyser_b200 import synth
from topdogspectrumanalyser_b200.sweep import gather_rows, max_shard, shard_bands


def test_shard_bands_matches_survey_split():
    sizes = [len(shard_bands(300, 8, r)) for r in range(8)]
    assert sizes == [38, 38, 38, 38, 37, 37, 37, 37]
    got = [b for r in range(8) for b in shard_bands(300, 8, r)]
    assert got == list(range(300))
    assert max_shard(300, 8) == 38
    assert [len(shard_bands(5, 8, r)) for r in range(8)] == [1, 1, 1, 1, 1, 0, 0, 0]      # ragged / empty shards
    assert list(shard_bands(7, 1, 0)) == list(range(7))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_bands, width, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_bands(n_bands, world, rank)
    local = torch.tensor([[float(b) * 1000 + k for k in range(width)] for b in mine], dtype=torch.float32).reshape(len(mine), width)
    rows = gather_rows(local, n_bands)
    q.put((rank, rows.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_bands", [7, 2, 1])
def test_gather_rows_gloo_world2(n_bands):
    world, width = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_bands, width, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.array([[b * 1000 + k for k in range(width)] for b in range(n_bands)], dtype=np.float32).reshape(n_bands, width)
    for r in range(world):
        np.testing.assert_array_equal(results[r], want)


def test_synth_shards_agree_with_full_array():
    full = synth.cfg4_subbands(n_bands=5, frames=2, n=64, seed=3)
    part = synth.cfg4_subbands(n_bands=5, frames=2, n=64, seed=3, bands=range(2, 4))
    np.testing.assert_array_equal(full[2:4], part)
    a = synth.cfg2_frames(b=4, n=128, seed=1)
    b = synth.cfg2_frames(b=4, n=128, seed=1)
    np.testing.assert_array_equal(a, b)
    assert a.dtype == np.complex64 and a.flags.c_contiguous


def test_datasource_interface_without_gpu():
    """The boundary class constructs and answers the 'not running' path without a device; start() fails loudly."""
    from topdogspectrumanalyser_b200.datasources import B200SampleDataSource, SampleDataSource, SyntheticIQFeed
    src = B200SampleDataSource(2048000, 98000000, feed=SyntheticIQFeed())
    assert isinstance(src, SampleDataSource)
    p, bins = src.get_power_levels()
    assert p.shape == bins.shape == (1024,) and not p.any()
    assert bins[0] == 98e6 - 1.024e6 and bins[-1] == 98e6 + 1.024e6
    src.set_averaging("exp", 8)
    src.reset_averaging()
    src.set_psd_mode(True)
    src.sample_count = 4096
    assert src.sample_count == 4096 and src.fft_size == 4096
    for name in ("start", "stop", "get_power_levels", "update_frequency", "update_centre_frequency",
                 "set_window_type", "set_psd_mode", "set_averaging", "reset_averaging", "read_samples_only",
                 "get_raw_samples", "pause", "resume", "set_gain"):
        assert callable(getattr(src, name)), name
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            src.start()


def test_sweep_csv_and_binary_parsers(golden):
    """Wire formats -> rows: same values the reference's _parse extracts (hackrf_sweep.py:138-146)."""
    from topdogspectrumanalyser_b200 import sweep_io
    g = golden("sweep_stitch.npz")
    text = g["csv_text"].tobytes()
    lo, hi, vals, nb, used = sweep_io.parse_csv(text, max_rows=64, max_bins=64)
    assert used == len(text)
    assert len(lo) == len(g["lo"])                      # the two malformed lines are skipped
    np.testing.assert_array_equal(lo, g["lo"].astype(np.float64))
    np.testing.assert_array_equal(hi, g["hi"].astype(np.float64))
    assert set(nb.tolist()) == {50}
    np.testing.assert_array_equal(vals[:, :50], g["rows"])
    # an incomplete last line is left for the next call
    cut = text[:len(text) - 7]
    lo2, _, _, _, used2 = sweep_io.parse_csv(cut, max_rows=64, max_bins=64)
    assert len(lo2) == len(lo) - 1 and cut[used2 - 1:used2] == b"\n"
    # binary -B records
    blob = g["binary"].tobytes()
    lo3, hi3, vals3, nb3, used3 = sweep_io.parse_binary(blob, max_rows=64, max_bins=64)
    assert used3 == len(blob)
    np.testing.assert_array_equal(lo3, lo)
    np.testing.assert_array_equal(hi3, hi)
    np.testing.assert_array_equal(vals3[:, :50], g["rows"])
    lo4, _, _, _, used4 = sweep_io.parse_binary(blob[:-3], max_rows=64, max_bins=64)
    assert len(lo4) == len(lo) - 1 and used4 < len(blob) - 3
    # rows wider than max_bins are dropped, never truncated silently
    lo5, _, _, _, _ = sweep_io.parse_csv(text, max_rows=64, max_bins=10)
    assert len(lo5) == 0


def test_hackrf_chunk_feed_consume_policy(golden):
    """Freshest-tail consume policy vs the executed HackrfSamplesDataSource._consume_samples (hackrf_samples.py:254-305)."""
    from topdogspectrumanalyser_b200.datasources import HackrfChunkFeed
    g = golden("hackrf_chain.npz")
    feed = HackrfChunkFeed(20e6, 2450e6, timeout=0.05)
    heads = iter(g["consume_heads"])
    for kind, arg in g["consume_script"]:
        if kind == 0:
            feed.put((np.arange(65536, dtype=np.float32) + 100000.0 * arg).astype(np.complex64))
        else:
            r = feed.read_samples(int(arg))
            first, last, n = next(heads)
            if n == 0:
                assert r is None
            else:
                assert len(r) == n and r[0].real == first and r[-1].real == last
    # drop-oldest on overflow (hackrf_samples.py:221-237)
    feed = HackrfChunkFeed(20e6, 2450e6)
    for k in range(6):
        feed.put(np.full(8, k, dtype=np.complex64))
    assert feed.stats["queue_overflows"] == 2 and feed.stats["samples_dropped"] == 16
    assert feed.read_samples(4)[0].real == 5.0


def test_sweep_source_interface_without_gpu():
    """The IQ-in sweep source builds the reference's grid (hackrf_sweep.py:32-40) and answers get_data without a device."""
    from topdogspectrumanalyser_b200.datasources import B200SweepDataSource, SweepDataSource, SyntheticTunerFeed
    src = B200SweepDataSource(88e6, 108e6, 30000, feed=SyntheticTunerFeed(20e6))
    assert isinstance(src, SweepDataSource)
    n = int((108e6 - 88e6) / 30000)
    np.testing.assert_array_equal(src.frequency_grid, np.linspace(88e6, 108e6, n))
    d = src.get_data()
    assert d.shape == (n,) and np.isnan(d).all() and src.get_number_of_points() == n
    assert src.n_bands == 1 and src.n_fft == 1024 and (src.lna_gain, src.vga_gain, src.amp_enabled) == (20, 20, True)
    src.set_gains(lna_gain=24)
    assert src.lna_gain == 24
    for name in ("start", "stop", "get_data", "get_number_of_points", "set_gains", "set_amplifier"):
        assert callable(getattr(src, name))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            src.start()
        assert not src.is_running
