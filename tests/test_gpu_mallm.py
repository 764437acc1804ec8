"""MA-LLM compressors on the GPU (rtk_mallm_compress) against the reference's stock torch op sequence executed on the
same GPU (oracle/reference_ops.py: the bit-exact target) and against the explicit CPU oracle (oracle/mallm.py)."""
import pytest
import torch

from helpers import scene_video

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vc():
    from retake import visual_compression
    return visual_compression


def _bank(seed, T, N, C, dup_every=5):
    g = torch.Generator().manual_seed(seed)
    return scene_video(g, T, N, C, dup_every=dup_every).to(torch.bfloat16)[None]


CASES = [  # T, N, C, t
    (24, 96, 256, 9),          # N < 128: strided ATen mean
    (20, 136, 512, 11),        # block width 16 in the mean
    (16, 256, 1152, 8),
    (14, 729, 256, 5),         # N % 8 != 0: position-dependent alignment of the mean
    (40, 64, 3584, 13),
    (33, 200, 264, 1),         # down to a single frame
    (9, 130, 256, 9),          # nothing to do
]


@pytest.mark.parametrize("T,N,C,t", CASES)
@pytest.mark.parametrize("sync", [False, True])
@pytest.mark.parametrize("hard", [False, True])
def test_fused_loop_matches_torch_cuda_ops(vc, T, N, C, t, sync, hard):
    from oracle import reference_ops as ro
    x = _bank(T * 1000 + N + int(sync) * 7 + int(hard) * 13, T, N, C).cuda()
    keep = x.clone()
    want, want_size = ro.mallm_compress(x.clone(), t, sync, hard)
    got, got_size = vc.mallm_compress(x, t, sync=sync, hard=hard)
    assert torch.equal(x, keep), "the input bank must not be modified"
    assert got.shape == (1, t, N, C) and torch.equal(got, want)
    if hard:
        assert got_size is None
    else:
        assert torch.equal(got_size, want_size)


@pytest.mark.parametrize("sync", [False, True])
def test_single_round_functions_chain_like_the_reference(vc, sync):
    """The reference's own call pattern (qwen2_vl.py:402-409): one function call per removed frame, sizes fed back in."""
    from oracle import reference_ops as ro
    T, N, C, t = 18, 160, 512, 6
    x = _bank(77, T, N, C).cuda()
    b, s = x, torch.ones_like(x[:, :, :, 0])
    rb, rs = x.clone(), torch.ones_like(x[:, :, :, 0])
    h, rh = x, x.clone()
    while b.shape[1] > t:
        b, s = vc.memory_bank_compress_MALLM(b, s, sync=sync)
        rb, rs = ro.mallm_round(rb, rs, sync)
        h = vc.memory_bank_compress_MALLM_hard(h, sync=sync)
        rh = ro.mallm_hard_round(rh, sync)
        assert torch.equal(b, rb) and torch.equal(s, rs) and torch.equal(h, rh)
    fused, fused_size = vc.mallm_compress(x, t, sync=sync)
    assert torch.equal(fused, b) and torch.equal(fused_size, s)


def test_large_sizes_are_counted_in_bf16_like_the_reference(vc):
    """Sizes live in bf16 (odd counts above 256 are not representable): 298 merges of identical frames."""
    from oracle import reference_ops as ro
    T, N, C = 300, 8, 256
    x = _bank(5, 4, N, C)[:, :1].expand(1, T, N, C).contiguous().cuda()          # identical frames: always merge at 0
    want, want_size = ro.mallm_compress(x.clone(), 2, False, False)
    got, got_size = vc.mallm_compress(x, 2)
    assert torch.equal(got, want) and torch.equal(got_size, want_size)
    assert float(got_size.max()) > 128.0


@pytest.mark.parametrize("sync", [False, True])
@pytest.mark.parametrize("hard", [False, True])
def test_explicit_cpu_oracle(vc, sync, hard):
    from oracle import mallm
    T, N, C, t = 12, 137, 256, 5
    x = _bank(11 + int(sync) + 2 * int(hard), T, N, C)
    want, want_size = mallm.mallm_compress(x, t, sync, hard, reduce="aten_cuda")
    got, got_size = vc.mallm_compress(x.cuda(), t, sync=sync, hard=hard)
    assert torch.equal(got.cpu(), want)
    if not hard:
        assert torch.equal(got_size.cpu(), want_size)


def test_full_size_qwen_shape_properties(vc):
    """Qwen2-VL 7B shape, 256 frames -> T=128, half kept: size conservation, untouched frames stay bit-identical."""
    T, N, C, t = 128, 256, 3584, 64
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(1, T, N, C, generator=g, device="cuda", dtype=torch.float32).to(torch.bfloat16)
    out, size = vc.mallm_compress(x, t)
    assert out.shape == (1, t, N, C) and size.shape == (1, t, N)
    assert torch.equal(size.float().sum(1), torch.full((1, N), float(T), device="cuda"))       # every frame counted once
    hard, _ = vc.mallm_compress(x, t, hard=True)
    # every surviving row of the hard variant is an input row of the same patch, in increasing frame order
    xin = x[0].float()
    match = (hard[0].float()[:, None] == xin[None]).all(-1)                                      # [t, T, N]
    assert bool(match.any(1).all())
    first = match.float().argmax(1)
    assert bool((first[1:] > first[:-1]).all())


def test_argument_errors(vc):
    x = torch.zeros(1, 4, 8, 256, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(ValueError):
        vc.mallm_compress(x[0], 2)
    with pytest.raises(ValueError):
        vc.mallm_compress(x, 0)
    with pytest.raises((TypeError, ValueError)):
        vc.mallm_compress(x.float(), 2)
    with pytest.raises(ValueError):
        vc.memory_bank_compress_MALLM(x[:, :1], torch.ones(1, 1, 8, dtype=torch.bfloat16, device="cuda"))
