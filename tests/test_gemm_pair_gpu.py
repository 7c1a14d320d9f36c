"""CTA-pair GEMM (csrc/gemm_pair.cu, tcgen05.mma.cta_group::2) against the one-CTA kernel and a plain
PyTorch fp32 reference of the same op.  Both kernels form the same products and accumulate each output
element over k in the same order, so their results must be BIT-IDENTICAL; against fp32 PyTorch the
tolerance is half an ulp of the 16-bit output format (stated per assertion)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rand(shape, dtype, seed, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(shape, generator=g, device=DEV) * scale).to(dtype)


@pytest.fixture
def pair_switch():
    from emdr2_b200 import ops
    before = ops.get_option("gemm_pair")
    yield ops
    ops.set_option("gemm_pair", before)


def _both(ops, fn):
    ops.set_option("gemm_pair", 0)
    one = fn()
    ops.set_option("gemm_pair", 1)
    two = fn()
    torch.cuda.synchronize()
    return one, two


# m, n, k, bias, gelu, residual — every shape has >= 148 tiles of 256 x 256, the pair kernel's threshold
CASES = [
    (7680, 1536, 768, False, False, False),
    (7700, 2304, 768, True, False, False),       # ragged M: the last pair's second CTA is past the end
    (16384, 768, 768, True, False, True),        # attention output projection + residual
    (12800, 3072, 768, True, True, False),       # h -> 4h + GeLU
    (12800, 768, 3072, True, False, True),       # 4h -> h + residual (48 k blocks: the ring wraps 8 times)
    (4100, 30720, 768, True, False, False),      # LM head shape, ragged M
    (20000, 520, 200, True, True, True),         # ragged N and K (partial boxes), all epilogues
]


@pytest.mark.parametrize("m,n,k,bias,gelu,res", CASES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_pair_gemm_equals_single_cta_kernel_and_fp32_reference(pair_switch, m, n, k, bias, gelu, res, dtype):
    ops = pair_switch
    x, w = _rand((m, k), dtype, 1), _rand((n, k), dtype, 2, scale=k ** -0.5)
    b = _rand((n,), dtype, 3) if bias else None
    r = _rand((m, n), dtype, 4) if res else None
    one, two = _both(ops, lambda: ops.linear(x, w, bias=b, gelu=gelu, residual=r))
    assert torch.equal(one, two), (one.float() - two.float()).abs().max().item()
    rows = torch.cat([torch.arange(0, 300), torch.arange(m - 300, m)]).to(DEV)      # reference on a slice
    want = x[rows].float() @ w.float().T
    if bias:
        want = want + b.float()
    if gelu:
        want = torch.nn.functional.gelu(want)
    if res:
        want = want + r[rows].float()
    tol = 2 ** -8 if dtype == torch.bfloat16 else 2 ** -11
    assert torch.allclose(two[rows].float(), want, rtol=tol, atol=tol), (two[rows].float() - want).abs().max().item()


def test_pair_gemm_training_epilogues_and_strided_views(pair_switch):
    """PREACT side output, GeLU-backward aux and row-strided operand / output views."""
    ops = pair_switch
    dtype = torch.bfloat16
    m, n, k = 9600, 1024, 768
    big = _rand((m, 2 * k), dtype, 5)
    x = big[:, k:]                                   # row pitch 2k
    w = _rand((n, k), dtype, 6, scale=k ** -0.5)
    b = _rand((n,), dtype, 7)

    def fwd():
        pre = torch.empty((m, n), dtype=dtype, device=DEV)
        outbuf = torch.zeros((m, n + 64), dtype=dtype, device=DEV)
        ops.gemm_ex(x, w, out=outbuf[:, 32:32 + n], bias=b, gelu=True, preact_out=pre)
        return torch.cat([outbuf, pre], dim=1)

    one, two = _both(ops, fwd)
    assert torch.equal(one, two)
    assert (two[:, :32] == 0).all() and (two[:, 32 + n:n + 64] == 0).all()
    u = _rand((m, n), dtype, 8)
    dy = _rand((m, k), dtype, 9)
    wt = _rand((n, k), dtype, 10, scale=k ** -0.5)   # dA[m, n] = dY[m, k] . Wt[n, k]^T, times GeLU'(u)
    one, two = _both(ops, lambda: ops.gemm_ex(dy, wt, gelu_bwd_aux=u))
    assert torch.equal(one, two)


def test_small_products_stay_on_the_single_cta_kernel(pair_switch):
    ops = pair_switch
    x, w = _rand((512, 768), torch.float16, 11), _rand((768, 768), torch.float16, 12, scale=0.03)
    one, two = _both(ops, lambda: ops.linear(x, w))
    assert torch.equal(one, two)
