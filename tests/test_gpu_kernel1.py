"""Parity of the fused window+FFT+PSD+dB kernel against the float64 oracle (through the C ABI)."""
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


def run_plan(dev, iq, n, window="hanning", mode="power", precision="f64", fs=2.048e6):
    import torch
    from topdogspectrumanalyser_b200.engine import SpectrumPlan
    plan = SpectrumPlan(n, window, mode=mode, precision=precision, fs=fs, device=dev)
    got = plan.psd_db(torch.from_numpy(np.ascontiguousarray(iq)).to(dev)).cpu().numpy()
    plan.close()
    assert got.dtype == np.float32
    return got.astype(np.float64)


def f32_tail_ok(got, want, near_tol=2e-3):
    """float32 butterflies: the error is a fixed absolute level in |X|, so only deep nulls move.

    Bins within 40 dB of the frame's mean bin must be within 1e-4 dB... (that is the same
    tail numpy's own float32 pocketfft shows: BASELINE.md section 2)."""
    err = np.abs(got - want)
    lin = 10 ** (want / 10)
    rel_db = want - 10 * np.log10(lin.mean(axis=1, keepdims=True))
    near = rel_db > -40.0
    assert np.median(err) < 2e-5, np.median(err)
    assert err[near].max() <= near_tol, err[near].max()
    assert (err > TOL_DB).mean() < 2e-2
    return err


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536])
def test_sizes_f64_every_bin(dev, n):
    b = 6 if n <= 8192 else 3
    iq = synth.cfg2_frames(b=b, n=n, seed=100 + n)
    want = O.power_db_batch(iq, O.make_window("hanning", n))
    got = run_plan(dev, iq, n)
    err = np.abs(got - want)
    assert err.max() <= TOL_DB, (n, err.max(), np.unravel_index(err.argmax(), err.shape))


@pytest.mark.parametrize("n", [64, 512, 1024, 2048, 4096, 8192, 16384, 65536])
def test_sizes_f32_fast_path(dev, n):
    iq = synth.cfg2_frames(b=4, n=n, seed=200 + n, tones=False)
    want = O.power_db_batch(iq, O.make_window("hanning", n))
    got = run_plan(dev, iq, n, precision="f32")
    f32_tail_ok(got, want)


@pytest.mark.parametrize("window", ["hanning", "hamming", "rectangle", "blackman"])
@pytest.mark.parametrize("mode", ["power", "psd", "mag20"])
def test_windows_and_modes(dev, window, mode):
    n = 2048
    iq = synth.cfg2_frames(b=4, n=n, seed=7)
    want = O.power_db_batch(iq, O.make_window(window, n), mode, fs=2.048e6)
    got = run_plan(dev, iq, n, window=window, mode=mode)
    assert np.abs(got - want).max() <= TOL_DB


def test_golden_reference_rows(dev, golden):
    """Device output against arrays the executed reference produced (tests/golden/rtl_chain.npz)."""
    g = golden("rtl_chain.npz")
    fs = float(g["fs"])
    got = run_plan(dev, g["cfg1_iq"], 1024, fs=fs)
    assert np.abs(got - g["cfg1_db"]).max() <= TOL_DB
    for w in ("hanning", "hamming", "rectangle"):
        for mode in ("power", "psd"):
            got = run_plan(dev, g["w_iq"], 4096, window=w, mode=mode, fs=fs)
            assert np.abs(got - g[f"w_{w}_{mode}"]).max() <= TOL_DB, (w, mode)
    for n in (512, 2048, 8192):
        got = run_plan(dev, g[f"n{n}_iq"], n, fs=fs)
        assert np.abs(got - g[f"n{n}_db"]).max() <= TOL_DB, n


def test_known_answers(dev, golden):
    g = golden("rtl_chain.npz")
    got = run_plan(dev, g["kat_iq"], 1024)
    assert np.abs(got - g["kat_db"]).max() <= TOL_DB
    np.testing.assert_allclose(got[0], -100.0, atol=TOL_DB)     # impulse at n=0: w[0] = 0
    np.testing.assert_allclose(got[1], -100.0, atol=TOL_DB)     # zeros
    assert int(np.argmax(got[3])) == 512 + 100                  # fftshift placement


def test_builtin_window_matches_numpy(dev):
    from topdogspectrumanalyser_b200.engine import SpectrumPlan, numpy_window
    for name in ("hanning", "hamming", "blackman", "rectangle"):
        plan = SpectrumPlan(1024, name, device=dev)
        plan.set_window_builtin(name)
        np.testing.assert_allclose(plan.window_table(), numpy_window(name, 1024), rtol=0, atol=4e-16)
        plan.close()
    plan = SpectrumPlan(1024, "hanning", device=dev)
    plan.set_window_builtin("hanning", "rms")
    np.testing.assert_allclose(plan.window_table(), numpy_window("hanning", 1024, "rms"), rtol=3e-7)
    plan.close()


def test_empty_and_strided(dev):
    import torch
    from topdogspectrumanalyser_b200.engine import SpectrumPlan
    plan = SpectrumPlan(1024, device=dev)
    out = plan.psd_db(torch.empty((0, 1024), dtype=torch.complex64, device=dev))
    assert out.shape == (0, 1024)
    # overlapping frames from a flat stream (frame_stride < N)
    stream = synth.cfg3_stream(n_samples=8192, seed=5)
    x = torch.from_numpy(stream).to(dev)
    got = plan.psd_db(x, n_frames=13, frame_stride=512).cpu().numpy().astype(np.float64)
    frames = np.stack([stream[i * 512:i * 512 + 1024] for i in range(13)])
    want = O.power_db_batch(frames, O.make_window("hanning", 1024))
    assert np.abs(got - want).max() <= TOL_DB
    with pytest.raises(ValueError):
        plan.psd_db(x, n_frames=20, frame_stride=512)
    plan.close()


def test_unaligned_frames_take_the_direct_load_path(dev):
    """Odd frame_stride / 8-byte-aligned base: the bulk-copy staging needs 16-byte frames, so the launcher must fall
    back to direct loads and still match (both precisions)."""
    import torch
    from topdogspectrumanalyser_b200.engine import SpectrumPlan
    stream = synth.cfg3_stream(n_samples=40000, seed=9)
    x = torch.from_numpy(stream).to(dev)
    for prec, tol in (("f64", TOL_DB), ("f32", 5e-2)):
        plan = SpectrumPlan(4096, precision=prec, device=dev)
        got = plan.psd_db(x[1:], n_frames=9, frame_stride=3001).cpu().numpy().astype(np.float64)   # base + 8 B, odd stride
        frames = np.stack([stream[1 + i * 3001:1 + i * 3001 + 4096] for i in range(9)])
        want = O.power_db_batch(frames, O.make_window("hanning", 4096))
        assert np.abs(got - want).max() <= tol, prec
        plan.close()


def test_many_more_frames_than_resident_ctas(dev):
    """Ring refills and the persistent frame loop: 3001 frames of 1024 points (odd count, > one wave), f64 and f32."""
    import torch
    from topdogspectrumanalyser_b200.engine import SpectrumPlan
    iq = synth.cfg2_frames(b=3001, n=1024, seed=12)
    want = O.power_db_batch(iq, O.make_window("hanning", 1024), workers=-1)
    x = torch.from_numpy(iq).to(dev)
    plan = SpectrumPlan(1024, device=dev)
    assert np.abs(plan.psd_db(x).cpu().numpy() - want).max() <= TOL_DB
    plan.set_precision("f32")
    f32_tail_ok(plan.psd_db(x).cpu().numpy().astype(np.float64), want)
    plan.close()


@pytest.mark.parametrize("n", [1 << 17, 1 << 20])
def test_largest_sizes_two_pass_head(dev, n):
    """N = 256*M path (two-pass head + M-point tails), up to the maximum supported size 2^20, both precisions."""
    iq = synth.cfg2_frames(b=2, n=n, seed=300)
    want = O.power_db_batch(iq, O.make_window("hanning", n))
    got = run_plan(dev, iq, n)
    assert np.abs(got - want).max() <= TOL_DB, n
    # float32 rounding error grows with the number of passes: 3.4e-3 dB on near-mean bins at 2^20 (5 passes)
    f32_tail_ok(run_plan(dev, iq, n, precision="f32"), want, near_tol=5e-3)


def test_nan_inf_and_extreme_inputs(dev):
    """A NaN or Inf sample poisons its whole frame (every bin depends on every sample) and nothing else: NaN -> all
    NaN; Inf -> every bin non-finite (which bins are Inf and which NaN depends on the butterfly order, in pocketfft
    too). Tiny and huge but finite magnitudes keep 1e-4 dB (float64 path)."""
    n = 1024
    iq = synth.cfg2_frames(b=6, n=n, seed=301)
    iq[1, 17] = np.nan
    iq[3, 900] = np.inf
    iq[4] *= np.float32(1e-18)                 # |X|^2 ~ 1e-33: far below the 1e-10 floor -> -100 dB row
    iq[5] *= np.float32(1e12)
    with np.errstate(all="ignore"):
        want = O.power_db_batch(iq, O.make_window("hanning", n))
    got = run_plan(dev, iq, n)
    assert np.isnan(got[1]).all() and np.isnan(want[1]).all()
    assert not np.isfinite(got[3]).any() and not np.isfinite(want[3]).any()
    for r in (0, 2, 4, 5):
        assert np.abs(got[r] - want[r]).max() <= TOL_DB, r
    assert np.abs(got[4] + 100.0).max() < 1e-3


def test_unsupported_size_fails_loudly(dev):
    from topdogspectrumanalyser_b200 import _lib
    from topdogspectrumanalyser_b200.engine import SpectrumPlan
    with pytest.raises(_lib.TdsaError):
        SpectrumPlan(1000, device=dev)


def test_cfg2_full_size_f64(dev):
    """BASELINE.json config 2 at full size: 8192 frames x 4096, every bin vs the oracle."""
    iq = synth.cfg2_frames(b=8192, n=4096, seed=1)
    got = run_plan(dev, iq, 4096)
    want = O.power_db_batch(iq, O.make_window("hanning", 4096), workers=-1)
    err = np.abs(got - want)
    assert err.max() <= TOL_DB, err.max()
    # size-independent property: Parseval, sum_k |X|^2 = N * sum_n |x w|^2 (floor removed)
    lin = 10 ** (got[:64] / 10) - 1e-10
    w = O.make_window("hanning", 4096)
    rhs = 4096 * np.sum(np.abs(iq[:64].astype(np.complex128) * w) ** 2, axis=1)
    np.testing.assert_allclose(lin.sum(axis=1), rhs, rtol=1e-5)
