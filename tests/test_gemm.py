"""`vnl_gemm_tf32` / `vnl_gemm_tf32_ex` (include/vnl_train.h) through the C ABI: the dense contractions of the PPO update
(reference: the `linen.Dense` matmuls inside `value_and_grad(compute_ppo_intention_loss)`, ppo_imitation/train.py:251-268)
against float64 torch matmuls of the same operands.  tests/test_learner.py only reaches the small tiles (480-row minibatches); this file
covers every tile shape the full-size update uses -- 128 x {64, 128, 256} and 256 x 256 (two accumulators in tensor memory) --, both
operand majors, ragged edges, split-K with its 16-byte vector reductions, and the fused swish / swish' epilogues.
Tolerances: TF32 (one tensor-core pass, 10-bit mantissa operands) 5e-3 of the largest entry; 3xTF32 1e-4 (the accumulator truncates:
chains of <= 8 K blocks); fused epilogues against the unfused result 1e-6 relative."""
import pytest

from conftest import pkg

pytestmark = pytest.mark.gpu


def _operands(M, N, K, a_mn, b_mn, seed):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    pad = lambda n: (n + 3) // 4 * 4
    A = torch.randn((K, pad(M)) if a_mn else (M, pad(K)), device="cuda", generator=g)
    B = torch.randn((K, pad(N)) if b_mn else (N, pad(K)), device="cuda", generator=g)
    bias = torch.randn(N, device="cuda", generator=g)
    Ad = (A[:, :M].T if a_mn else A[:, :K]).double()
    Bd = (B[:, :N].T if b_mn else B[:, :K]).double()
    return A, B, bias, Ad @ Bd.T


# (M, N, K, splitk): which tile the launcher picks is stated per case
SHAPES = [
    (300, 60, 256, 1),      # 128 x 64, ragged in M and N
    (777, 1024, 232, 1),    # 128 x 128 (7 x 4 = 28 tiles of 256 would not fill the machine), ragged M, K not a multiple of 32
    (5000, 512, 96, 1),     # 128 x 128: 40 x 2 tiles of 128 x 256 = 80 would not fill the machine
    (18944, 256, 64, 1),    # 128 x 256: exactly 148 tiles, one wave
    (5376, 1024, 232, 1),   # 256 x 256: 42 x 4 = 168 tiles of 128 x 256 need a second wave -> 21 x 4 big tiles (value MLP layer 0 forward)
    (5120, 1024, 1024, 1),  # 256 x 256 (value MLP layer 1 forward / dgrad)
    (1024, 1024, 5120, 9),  # 256 x 256 with split-K 9: 16 x 9 = 144 CTAs (value MLP layer 1 wgrad), vector red.add
    (1000, 512, 5120, 12),  # 256 x 256 split-K, ragged M
    (1024, 1024, 5120, 2),  # 128 x 128 with split-K 2 (16 x 2 big CTAs would leave the machine empty)
    (256, 128, 5120, 40),   # policy wgrad: 2 tiles x 40 splits
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("majors", [(0, 1), (0, 0), (1, 1), (1, 0)])  # forward, dgrad, wgrad, (unused combination)
def test_gemm_all_tiles_and_majors(shape, majors):
    import torch
    tk = pkg("train_kernels")
    M, N, K, sk = shape
    a_mn, b_mn = majors
    A, B, bias, want = _operands(M, N, K, a_mn, b_mn, seed=M + 3 * N + 7 * K + a_mn + 2 * b_mn)
    want = want + bias.double()
    scale = float(want.abs().max())
    C = torch.full((M, (N + 3) // 4 * 4), 7.0, device="cuda")  # poisoned: every element of [:, :N] must be written (or zeroed + added)
    tk.gemm(A, a_mn, B, b_mn, C, M, N, K, bias=bias, splitk=sk)
    torch.cuda.synchronize()
    e1 = float((C[:, :N].double() - want).abs().max()) / scale
    assert e1 < 5e-3, e1
    C3 = torch.full_like(C, -3.0)
    tk.gemm(A, a_mn, B, b_mn, C3, M, N, K, bias=bias, x3=True, splitk=max(sk, (K + 255) // 256))
    torch.cuda.synchronize()
    e3 = float((C3[:, :N].double() - want).abs().max()) / scale
    assert e3 < 1e-4, e3


@pytest.mark.parametrize("shape", [(5376, 1024, 232), (640, 1024, 1024), (300, 256, 64)])  # big tile, 128 x 128, ragged M
def test_fused_swish_epilogues(shape):
    """epilogue 1 (forward): C = pre-activation, aux = swish(C); epilogue 2 (dgrad): C = (A . B^T) * swish'(aux)."""
    import torch
    tk = pkg("train_kernels")
    M, N, K = shape
    A, B, bias, _ = _operands(M, N, K, 0, 1, seed=11 + M)
    plain = torch.zeros(M, N, device="cuda")
    tk.gemm(A, 0, B, 1, plain, M, N, K, bias=bias)
    pre, act = torch.full((M, N), 5.0, device="cuda"), torch.full((M, N), 5.0, device="cuda")
    tk.gemm(A, 0, B, 1, pre, M, N, K, bias=bias, epilogue=1, aux=act)
    torch.cuda.synchronize()
    assert torch.equal(pre, plain)  # same tiles, same accumulation order: the pre-activation is bit-identical
    want = plain.double() * torch.sigmoid(plain.double())
    assert float((act.double() - want).abs().max()) < 1e-6 * max(1.0, float(want.abs().max()))
    # dgrad through the activation in front: A2 [M, N] . W[K2, N]^T with W K-major (N = in, contiguous `out`)
    g = torch.Generator(device="cuda").manual_seed(5)
    dy = torch.randn(M, N, device="cuda", generator=g)
    W = torch.randn(256, N, device="cuda", generator=g)  # flax kernel [in = 256, out = N]
    pre_in = torch.randn(M, 256, device="cuda", generator=g)
    d_plain = torch.zeros(M, 256, device="cuda")
    tk.gemm(dy, 0, W, 0, d_plain, M, 256, N)
    d_fused = torch.full((M, 256), 9.0, device="cuda")
    tk.gemm(dy, 0, W, 0, d_fused, M, 256, N, epilogue=2, aux=pre_in)
    torch.cuda.synchronize()
    s = torch.sigmoid(pre_in.double())
    want = d_plain.double() * (s + pre_in.double() * s * (1 - s))
    assert float((d_fused.double() - want).abs().max()) < 1e-6 * max(1.0, float(want.abs().max()))


def test_fused_epilogue_argument_checks():
    """A fused activation needs one pass over K and whole 16-byte row segments: anything else is refused (-5), not mis-computed."""
    import torch
    tk = pkg("train_kernels")
    A, B, bias, _ = _operands(256, 128, 512, 0, 1, seed=2)
    C, aux = torch.zeros(256, 128, device="cuda"), torch.zeros(256, 128, device="cuda")
    with pytest.raises(RuntimeError):
        tk.gemm(A, 0, B, 1, C, 256, 128, 512, bias=bias, splitk=2, epilogue=1, aux=aux)
    A2, B2, bias2, _ = _operands(256, 60, 512, 0, 1, seed=3)
    C2, aux2 = torch.zeros(256, 60, device="cuda"), torch.zeros(256, 60, device="cuda")
    with pytest.raises(RuntimeError):
        tk.gemm(A2, 0, B2, 1, C2, 256, 60, 512, bias=bias2, epilogue=1, aux=aux2)
