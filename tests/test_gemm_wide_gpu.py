"""Sixteen-warp ("wide") epilogue of the one-CTA GEMM (csrc/gemm.cu, kWide) against the eight-warp epilogue and a
plain PyTorch fp32 reference of the same op.  Both variants share the operand pipeline and apply the same fp32
epilogue arithmetic to each accumulator, so their results must be BIT-IDENTICAL; against fp32 PyTorch the tolerance
is half an ulp of the 16-bit output format (stated per assertion)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rand(shape, dtype, seed, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(shape, generator=g, device=DEV) * scale).to(dtype)


@pytest.fixture
def wide_switch():
    from emdr2_b200 import ops
    before = ops.get_option("gemm_wide"), ops.get_option("gemm_pair")
    ops.set_option("gemm_pair", 0)
    yield ops
    ops.set_option("gemm_wide", before[0])
    ops.set_option("gemm_pair", before[1])


def _both(ops, fn):
    ops.set_option("gemm_wide", 0)
    narrow = fn()
    ops.set_option("gemm_wide", 2)       # whenever eligible, also bias-only epilogues
    wide = fn()
    torch.cuda.synchronize()
    return narrow, wide


def test_default_mode_is_auto():
    from emdr2_b200 import ops
    assert ops.get_option("gemm_wide") == 1


# m, n, k, bias, gelu
CASES = [
    (12800, 3072, 768, True, True),        # h -> 4h + GeLU, the shape the variant exists for
    (7700, 2304, 768, True, False),        # ragged M, bias only
    (4100, 1536, 768, False, False),       # no epilogue arithmetic at all
    (20000, 520, 200, True, True),         # ragged N (the last 32-column box holds 8 columns) and ragged K
    (300, 40, 64, True, True),             # one partial tile: three column quarters have nothing to do
    (66000, 3072, 768, True, True),        # several tiles per CTA: the accumulator buffers alternate
]


@pytest.mark.parametrize("m,n,k,bias,gelu", CASES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_wide_epilogue_equals_narrow_and_fp32_reference(wide_switch, m, n, k, bias, gelu, dtype):
    ops = wide_switch
    x, w = _rand((m, k), dtype, 1), _rand((n, k), dtype, 2, scale=k ** -0.5)
    b = _rand((n,), dtype, 3) if bias else None
    narrow, wide = _both(ops, lambda: ops.linear(x, w, bias=b, gelu=gelu))
    assert torch.equal(narrow, wide), (narrow.float() - wide.float()).abs().max().item()
    rows = torch.cat([torch.arange(0, min(300, m)), torch.arange(max(m - 300, 0), m)]).to(DEV)
    want = x[rows].float() @ w.float().T
    if bias:
        want = want + b.float()
    if gelu:
        want = torch.nn.functional.gelu(want)
    tol = 2 ** -8 if dtype == torch.bfloat16 else 2 ** -11
    assert torch.allclose(wide[rows].float(), want, rtol=tol, atol=tol), (wide[rows].float() - want).abs().max().item()


def test_wide_epilogue_preact_output_strided_views_and_mn_major_operand(wide_switch):
    ops = wide_switch
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

    narrow, wide = _both(ops, fwd)
    assert torch.equal(narrow, wide)
    assert (wide[:, :32] == 0).all() and (wide[:, 32 + n:n + 64] == 0).all()      # nothing written beside the view
    pre = wide[:, n + 64:].float()
    want = x.float() @ w.float().T + b.float()
    assert torch.allclose(pre, want, rtol=2 ** -8, atol=2 ** -8)
    assert torch.allclose(wide[:, 32:32 + n].float(), torch.nn.functional.gelu(want), rtol=2 ** -8, atol=2 ** -8)
    # dX = dY . W with W [n, k] read in place as the [k', n'] operand
    dy = _rand((m, n), dtype, 8)
    narrow, wide = _both(ops, lambda: ops.gemm_ex(dy, w, b_mn=True))
    assert torch.equal(narrow, wide)
    assert torch.allclose(wide[:256].float(), dy[:256].float() @ w.float(), rtol=2 ** -8, atol=2 ** -6)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("m,n,k", [(9600, 3072, 768), (4100, 520, 200), (300, 40, 64)])
def test_wide_epilogue_with_an_aux_tile(wide_switch, m, n, k, dtype):
    """Residual and GeLU-backward epilogues (kEpi 2: three operand stages, the aux box of a quarter's next tile
    prefetched into its staging box a tile ahead): same arithmetic as the eight-warp variant, same bits."""
    ops = wide_switch
    x, w = _rand((m, k), dtype, 21), _rand((n, k), dtype, 22, scale=k ** -0.5)
    b = _rand((n,), dtype, 23)
    big = _rand((m, n + 24), dtype, 24)
    r = big[:, 16:16 + n]                                  # row pitch n + 24, 32-byte offset
    narrow, wide = _both(ops, lambda: ops.linear(x, w, bias=b, residual=r))
    assert torch.equal(narrow, wide)
    tol = 2 ** -8 if dtype == torch.bfloat16 else 2 ** -11
    want = x[:300].float() @ w.float().T + b.float() + r[:300].float()
    assert torch.allclose(wide[:300].float(), want, rtol=tol, atol=tol)
    # dU = (dA . W2) * GeLU'(u), W2 [k, n] read in place as the [k', n'] operand
    w2 = _rand((k, n), dtype, 25, scale=k ** -0.5)
    narrow, wide = _both(ops, lambda: ops.gemm_ex(x, w2, b_mn=True, gelu_bwd_aux=r))
    assert torch.equal(narrow, wide)
    u = r[:300].float().requires_grad_(True)
    torch.nn.functional.gelu(u).sum().backward()
    want = (x[:300].float() @ w2.float()) * u.grad
    assert torch.allclose(wide[:300].float(), want, rtol=4 * tol, atol=4 * tol)


def test_accumulating_products_stay_on_the_narrow_variant(wide_switch):
    """fp32-accumulating (split-K) products are not eligible: forcing the option changes nothing."""
    ops = wide_switch
    dtype = torch.float16
    x, r = _rand((2048, 768), dtype, 11), _rand((2048, 768), dtype, 13)

    def dw():
        acc = torch.zeros((768, 768), dtype=torch.float32, device=DEV)
        ops.gemm_ex(r, x, a_mn=True, b_mn=True, accumulate_into=acc, splits=1)
        return acc

    narrow, wide = _both(ops, dw)
    assert torch.equal(narrow, wide)
