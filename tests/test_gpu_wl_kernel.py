"""The warp-local 4096-point kernel (csrc/tdsa_fft_wl.cuh): tensor-map staging, team-local sub-transforms and the
dynamic frame scheduler, against the float64 oracle and against the classic kernel (TDSA_WL=0 is read once per
process, so the cross-check runs the classic kernel through an unaligned view, which always takes it)."""
import numpy as np
import pytest

from oracle import oracle as O
from topdogspectrumanalyser_b200 import synth

pytestmark = pytest.mark.gpu

TOL_DB = 1e-4
N = 4096


@pytest.fixture(scope="module")
def dev():
    import torch
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _plan(dev, **kw):
    from topdogspectrumanalyser_b200.engine import SpectrumPlan
    return SpectrumPlan(N, device=dev, **kw)


def test_launch_geometry_is_the_warp_local_kernel(dev):
    plan = _plan(dev)
    info = plan.info()
    plan.close()
    assert info["threads_per_cta"] == 256 and info["ctas_per_sm"] == 2, info
    assert info["smem_bytes"] > 100 * 1024, info       # exchange regions + swizzled stage(s), two CTAs per SM


@pytest.mark.parametrize("frames", [1, 2, 295, 296, 297, 1777])
def test_frame_counts_around_the_grid_size(dev, frames):
    """Fewer frames than CTAs, exactly one per CTA, one more, and several per CTA (odd count): every frame is claimed
    exactly once by the dynamic scheduler and lands in its own row."""
    import torch
    iq = synth.cfg2_frames(b=frames, n=N, seed=300 + frames)
    want = O.power_db_batch(iq, O.make_window("hanning", N), workers=-1)
    plan = _plan(dev)
    got = plan.psd_db(torch.from_numpy(iq).to(dev)).cpu().numpy().astype(np.float64)
    plan.close()
    assert np.abs(got - want).max() <= TOL_DB


def test_scheduler_rearms_between_launches(dev):
    """Back-to-back launches of different sizes on one plan: the frame counter must return to zero every time."""
    import torch
    iq = synth.cfg2_frames(b=700, n=N, seed=41)
    want = O.power_db_batch(iq, O.make_window("hanning", N), workers=-1)
    x = torch.from_numpy(iq).to(dev)
    plan = _plan(dev)
    for b in (700, 3, 512, 1, 700, 299):
        got = plan.psd_db(x[:b].contiguous()).cpu().numpy().astype(np.float64)
        assert got.shape == (b, N)
        assert np.abs(got - want[:b]).max() <= TOL_DB, b
    plan.close()


def test_overlapping_frames_through_the_tensor_map(dev):
    """frame_stride < N (Welch-style 50 % and 75 % overlap) and a stride larger than N: the 3-D tensor map's outer
    stride is the frame stride, whatever it is (it only has to be even)."""
    import torch
    stream = synth.cfg3_stream(n_samples=1 << 18, seed=17)
    x = torch.from_numpy(stream).to(dev)
    plan = _plan(dev)
    for stride, frames in ((2048, 100), (1024, 61), (5000, 40), (4098, 30)):
        got = plan.psd_db(x, n_frames=frames, frame_stride=stride).cpu().numpy().astype(np.float64)
        rows = np.stack([stream[i * stride:i * stride + N] for i in range(frames)])
        want = O.power_db_batch(rows, O.make_window("hanning", N))
        assert np.abs(got - want).max() <= TOL_DB, stride
    plan.close()


def test_matches_the_classic_kernel_bit_for_bit_in_float64(dev):
    """Same butterflies, same twiddle tables, same operation order per bin: the two schedules must agree to the bit.
    The classic kernel is reached through an 8-byte-aligned view of the same samples, which TMA cannot stage."""
    import torch
    frames = 63
    stream = synth.cfg3_stream(n_samples=frames * N + 2, seed=23)
    x = torch.from_numpy(stream).to(dev)
    plan = _plan(dev)
    classic = plan.psd_db(x[1:], n_frames=frames, frame_stride=N).cpu().numpy()          # base + 8 bytes
    copy = torch.from_numpy(np.ascontiguousarray(stream[1:1 + frames * N])).to(dev)        # aligned copy
    wl = plan.psd_db(copy.view(frames, N)).cpu().numpy()
    plan.close()
    rows = stream[1:1 + frames * N].reshape(frames, N)
    want = O.power_db_batch(rows, O.make_window("hanning", N))
    assert np.abs(wl.astype(np.float64) - want).max() <= TOL_DB
    assert np.array_equal(wl, classic), np.abs(wl - classic).max()


@pytest.mark.parametrize("window,mode", [("blackman", "psd"), ("hamming", "mag20"), ("rectangle", "power")])
def test_windows_and_modes_at_4096(dev, window, mode):
    import torch
    iq = synth.cfg2_frames(b=40, n=N, seed=77)
    want = O.power_db_batch(iq, O.make_window(window, N), mode, fs=2.048e6)
    plan = _plan(dev, window=window, mode=mode, fs=2.048e6)
    got = plan.psd_db(torch.from_numpy(iq).to(dev)).cpu().numpy().astype(np.float64)
    plan.close()
    assert np.abs(got - want).max() <= TOL_DB


def test_linear_epilogue_and_float32_path(dev):
    import torch
    iq = synth.cfg2_frames(b=333, n=N, seed=5, tones=False)
    w = O.make_window("hanning", N)
    x = torch.from_numpy(iq).to(dev)
    plan = _plan(dev)
    lin = plan.power_linear(x).cpu().numpy()
    want_lin = np.abs(np.fft.fftshift(np.fft.fft(iq.astype(np.complex128) * w, axis=1), axes=1)) ** 2
    assert np.abs(lin - want_lin).max() <= 1e-10 * want_lin.max()
    plan.set_precision("f32")
    got = plan.psd_db(x).cpu().numpy().astype(np.float64)
    plan.close()
    want = O.power_db_batch(iq, w, workers=-1)
    err = np.abs(got - want)
    rel_db = want - 10 * np.log10((10 ** (want / 10)).mean(axis=1, keepdims=True))
    assert np.median(err) < 2e-5 and err[rel_db > -40].max() <= 2e-3


def test_nan_and_inf_frames_do_not_leak_into_neighbours(dev):
    import torch
    iq = synth.cfg2_frames(b=600, n=N, seed=9)
    iq[17, 100] = np.nan
    iq[401, 7] = np.inf
    want = O.power_db_batch(iq, O.make_window("hanning", N), workers=-1)
    plan = _plan(dev)
    got = plan.psd_db(torch.from_numpy(iq).to(dev)).cpu().numpy().astype(np.float64)
    plan.close()
    bad = np.zeros(600, dtype=bool)
    bad[[17, 401]] = True
    assert np.isnan(got[17]).all() and not np.isfinite(got[401]).any()
    assert np.abs(got[~bad] - want[~bad]).max() <= TOL_DB
