"""tcgen05 kernels (bf16 GEMM with fused epilogues, cosine attention) against fp32/fp64 PyTorch math on the
same bf16-rounded operands.  Tolerances are the bf16 output rounding (2^-9 relative) unless stated."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2
from oracle import mapdit_oracle as O

pytestmark = pytest.mark.gpu


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


SHAPES = [(128, 32, 64), (256, 256, 64), (128, 64, 128), (512, 1152, 384), (200, 96, 384), (4096, 2304, 768), (8192, 768, 3072),
          (1000, 1536, 384), (256, 4608, 768), (384, 384, 1536), (5248, 1024, 128), (5200, 1024, 192), (16384, 1152, 384)]


@pytest.mark.parametrize("m,n,k", SHAPES)
@pytest.mark.parametrize("two_cta", [0, 1])
def test_gemm_bf16_store(m, n, k, two_cta):
    from mapdit_b200 import _lib, ops
    _lib.set_option("gemm_2cta", two_cta)  # 1 = cta_group::2 kernel where the shape qualifies
    a, b = rnd(m, k, seed=1).bfloat16(), rnd(n, k, seed=2, scale=k ** -0.5).bfloat16()
    ref = a.double() @ b.double().t()
    out32 = torch.full((m, n), float("nan"), device="cuda")
    ops.gemm_bf16(a, b, out32)
    assert rel_l2(out32, ref) < 1e-5, "fp32-out GEMM must only differ by accumulation order"
    out16 = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    ops.gemm_bf16(a, b, out16)
    assert rel_l2(out16.float(), ref) < 3e-3
    # strided A (column slice of a wider buffer), as the engine uses for q|k|v style views
    wide = torch.zeros(m, k + 64, device="cuda", dtype=torch.bfloat16)
    wide[:, 64:] = a
    ops.gemm_bf16(wide[:, 64:], b, out32)
    assert rel_l2(out32, ref) < 1e-5


# the last four shapes reach the cta_group::2 kernel also for the N = D GEMMs (second-generation residual epilogues): 256-wide tiles
# (N = 256, 768), 128-wide tiles with 64 tokens per sample (N = 384), and a clipped last 256-wide tile (N = 1152)
@pytest.mark.parametrize("N,T,D", [(2, 64, 384), (3, 256, 256), (1, 256, 768), (33, 64, 384), (40, 256, 256), (41, 128, 768), (160, 64, 384),
                                   (80, 256, 256), (30, 256, 768), (200, 64, 384), (16, 256, 1152)])
@pytest.mark.parametrize("two_cta", [0, 1, 2])
def test_gemm_bf16_fused_epilogues(N, T, D, two_cta):
    from mapdit_b200 import _lib, ops
    _lib.set_option("gemm_2cta", min(two_cta, 1))
    _lib.set_option("gemm_fused_resid", 0 if two_cta == 2 else 1)  # 2 = CTA-pair kernel with the first-generation residual epilogue
    try:
        _fused_epilogues(N, T, D)
    finally:
        _lib.set_option("gemm_fused_resid", 1)
        _lib.set_option("gemm_2cta", 1)


def _fused_epilogues(N, T, D):
    from mapdit_b200 import _lib, ops
    M, hd = N * T, 64
    h = rnd(M, D, seed=3).bfloat16()
    wqkv = rnd(3 * D, D, seed=4, scale=D ** -0.5).bfloat16()
    acc = h.float() @ wqkv.float().t()
    # QKNORM
    qkv = torch.empty(M, 3 * D, device="cuda", dtype=torch.bfloat16)
    ops.gemm_bf16(h, wqkv, qkv, epilogue=_lib.EPI_QKNORM, tokens=T, head_dim=hd, qk_cols=2 * D)
    ref = acc.clone()
    qk = ref[:, :2 * D].reshape(M, 2 * D // hd, hd)
    ref[:, :2 * D] = (qk * math.sqrt(hd) / (qk.norm(dim=-1, keepdim=True) + 1e-4)).reshape(M, 2 * D)
    assert rel_l2(qkv.float(), ref) < 3e-3
    # MPSILU (+ pre-activation copy)
    w1 = rnd(4 * D, D, seed=5, scale=D ** -0.5).bfloat16()
    z = h.float() @ w1.float().t()
    u = torch.empty(M, 4 * D, device="cuda", dtype=torch.bfloat16)
    pre = torch.empty_like(u)
    ops.gemm_bf16(h, w1, u, epilogue=_lib.EPI_MPSILU, out2=pre)
    assert rel_l2(u.float(), F.silu(z) / 0.596) < 3e-3
    assert rel_l2(pre.float(), z) < 3e-3
    # SILU_BWD: dgrad of fc2 fused with MPSiLU's backward, dz = (dy @ W2) * d/dz[silu(z)/0.596]
    dy = rnd(M, D, seed=11).bfloat16()
    w2t = rnd(4 * D, D, seed=12, scale=D ** -0.5).bfloat16()
    zb = pre.clone()
    dz = torch.empty(M, 4 * D, device="cuda", dtype=torch.bfloat16)
    ops.gemm_bf16(dy, w2t, dz, epilogue=_lib.EPI_SILU_BWD, resid=zb)
    zf = zb.float()
    sg = torch.sigmoid(zf)
    assert rel_l2(dz.float(), (dy.float() @ w2t.float().t()) * sg * (1 + zf * (1 - sg)) / 0.596) < 3e-3
    # RESID_MOD, in place on x
    x = rnd(M, D, seed=6).bfloat16()
    mods = rnd(N, 6 * D, seed=7)
    gain = torch.tensor(0.31, device="cuda")
    wo = rnd(D, D, seed=8, scale=D ** -0.5).bfloat16()
    y = h.float() @ wo.float().t()
    n_of_row = torch.arange(M, device="cuda") // T
    gate, shift, scale = mods[:, 2 * D:3 * D][n_of_row], mods[:, 3 * D:4 * D][n_of_row], mods[:, 4 * D:5 * D][n_of_row]
    xn = (x.float() + 0.3 * (gate * y - x.float())) / math.sqrt(0.7 ** 2 + 0.3 ** 2)
    g = float(gain)
    hn = ((xn * scale) + g * (shift - xn * scale)) / math.sqrt((1 - g) ** 2 + g ** 2)
    xio = x.clone()
    hout = torch.full_like(x, float("nan"))
    ops.gemm_bf16(h, wo, xio, epilogue=_lib.EPI_RESID_MOD, out2=hout, resid=xio, gate=mods[:, 2 * D:], shift=mods[:, 3 * D:],
                  scale=mods[:, 4 * D:], gain=gain, ldmod=mods.shape[1], tokens=T)
    assert rel_l2(xio.float(), xn) < 3e-3
    assert rel_l2(hout.float(), hn) < 3e-3
    # training flavour: separate output, raw branch output saved in `aux`; the same bits as the in-place eval call
    xo2, hout2, aux = (torch.full_like(x, float("nan")) for _ in range(3))
    ops.gemm_bf16(h, wo, xo2, epilogue=_lib.EPI_RESID_MOD, out2=hout2, resid=x, gate=mods[:, 2 * D:], shift=mods[:, 3 * D:],
                  scale=mods[:, 4 * D:], gain=gain, ldmod=mods.shape[1], tokens=T, aux=aux)
    assert torch.equal(xo2, xio) and torch.equal(hout2, hout)
    assert rel_l2(aux.float(), y) < 3e-3
    xio = x.clone()
    ops.gemm_bf16(h, wo, xio, epilogue=_lib.EPI_RESID, resid=xio, gate=mods[:, 2 * D:], ldmod=mods.shape[1], tokens=T)
    assert rel_l2(xio.float(), xn) < 3e-3
    xo2, aux = torch.full_like(x, float("nan")), torch.full_like(x, float("nan"))
    ops.gemm_bf16(h, wo, xo2, epilogue=_lib.EPI_RESID, resid=x, gate=mods[:, 2 * D:], ldmod=mods.shape[1], tokens=T, aux=aux)
    assert torch.equal(xo2, xio) and rel_l2(aux.float(), y) < 3e-3
    # RESID_ROT (rotation modulation, UNPINNED: SURVEY.md §A.8): h = R(rot * gain) x' (* scale), (cos, sin) table from rot_table
    rot = mods[:, 5 * D:5 * D + D // 2]
    cs = torch.full((N, 2 * D + 8), float("nan"), device="cuda")  # table in a wider buffer: column slice at offset 8, own ld
    ops.rot_table(rot, gain, cs[:, 8:], mods.shape[1], D, rot, gain, cs[:, 8 + D:])
    th = (rot * g).double()
    assert float((cs[:, 8:8 + D:2].double() - th.cos()).abs().max()) < 1e-6 and float((cs[:, 9:9 + D:2].double() - th.sin()).abs().max()) < 1e-6
    assert torch.equal(cs[:, 8:8 + D], cs[:, 8 + D:])
    thr = th.float()[n_of_row]
    xe, xo_ = xn[:, 0::2], xn[:, 1::2]
    rotated = torch.stack([xe * thr.cos() - xo_ * thr.sin(), xe * thr.sin() + xo_ * thr.cos()], dim=-1).reshape(M, D)
    for sc_arg, href in ((mods[:, 4 * D:], rotated * scale), (None, rotated)):
        xio = x.clone()
        hout = torch.full_like(x, float("nan"))
        ops.gemm_bf16(h, wo, xio, epilogue=_lib.EPI_RESID_ROT, out2=hout, resid=xio, gate=mods[:, 2 * D:], shift=cs[:, 8:], scale=sc_arg,
                      ldmod=mods.shape[1], ldrot=cs.stride(0), tokens=T)
        assert rel_l2(xio.float(), xn) < 3e-3
        assert rel_l2(hout.float(), href) < 3e-3
        xo2, hout2, aux = (torch.full_like(x, float("nan")) for _ in range(3))
        ops.gemm_bf16(h, wo, xo2, epilogue=_lib.EPI_RESID_ROT, out2=hout2, resid=x, gate=mods[:, 2 * D:], shift=cs[:, 8:], scale=sc_arg,
                      ldmod=mods.shape[1], ldrot=cs.stride(0), tokens=T, aux=aux)
        assert torch.equal(xo2, xio) and torch.equal(hout2, hout) and rel_l2(aux.float(), y) < 3e-3


@pytest.mark.parametrize("v2", [1, 0])
@pytest.mark.parametrize("N,T,H", [(2, 256, 6), (3, 64, 4), (1, 1024, 2), (5, 256, 12), (2, 128, 4), (40, 256, 12), (3, 512, 5)])
def test_cos_attn_bf16(N, T, H, v2):
    """both tcgen05 forward kernels (attn_v2: one CTA per SM, P in TMEM, ones-column row sum; v1: P through smem)"""
    from mapdit_b200 import _lib, ops
    hd, D = 64, H * 64
    qkv = rnd(N * T, 3 * D, seed=9)
    ops.qk_normalize(qkv, D, hd)
    qkv16 = qkv.bfloat16()
    q, k, v = qkv16.float().view(N, T, 3, H, hd).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q.double(), k.double(), v.double(), scale=1 / math.sqrt(hd)).transpose(1, 2).reshape(N * T, D)
    ref_lse = torch.logsumexp(q.double() @ k.double().transpose(-1, -2) / math.sqrt(hd), dim=-1).permute(0, 2, 1).reshape(N * T, H)
    o = torch.full((N * T, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    lse = torch.full((N * T, H), float("nan"), device="cuda")
    _lib.set_option("attn_v2", v2)
    try:
        ops.cos_attn(qkv16, o, N, T, H, hd, lse=lse)
        torch.cuda.synchronize()
    finally:
        _lib.set_option("attn_v2", 1)
    # P is rounded to bf16 before the PV product on the tensor-core path
    assert rel_l2(o.float(), ref) < 6e-3
    assert float((lse.double() - ref_lse).abs().max()) < 2e-2


@pytest.mark.parametrize("N,T,H", [(2, 256, 4), (30, 256, 12), (80, 256, 4), (16, 256, 18), (3, 64, 6)])
@pytest.mark.parametrize("two_cta", [0, 1])
def test_gemm_store_delta_epilogue_and_attention_backward_without_o(N, T, H, two_cta):
    """MAPDIT_EPI_STORE_DELTA: the out-proj dgrad GEMM also emits delta[row, head] = dO.O (autograd of SDPA,
    src/layers/attention.py:47), and the fused attention backward accepts it in place of o (o = NULL)"""
    from mapdit_b200 import _lib, ops
    D, M, hd = H * 64, N * T, 64
    _lib.set_option("gemm_2cta", two_cta)
    try:
        dy, wt = rnd(M, D, seed=21).bfloat16(), rnd(D, D, seed=22, scale=D ** -0.5).bfloat16()
        o = rnd(M, D, seed=23).bfloat16()
        plain = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
        ops.gemm_bf16(dy, wt, plain)
        dh = torch.full_like(plain, float("nan"))
        delta = torch.full((M, H), float("nan"), device="cuda")
        ops.gemm_bf16(dy, wt, dh, epilogue=_lib.EPI_STORE_DELTA, resid=o, aux=delta)
        assert torch.equal(dh, plain)
        ref = (plain.double() * o.double()).view(M, H, hd).sum(-1)
        assert rel_l2(delta, ref) < 1e-5
    finally:
        _lib.set_option("gemm_2cta", 1)
    if T != 256:
        return
    qkv = rnd(M, 3 * D, seed=24)
    sc = torch.empty(M, 2 * H, device="cuda")
    ops.qk_normalize_save(qkv, sc, D, hd)
    qkv = qkv.bfloat16()
    att = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(M, H, device="cuda")
    ops.cos_attn(qkv, att, N, T, H, hd, lse=lse)
    d_ref, d_new = torch.full_like(qkv, float("nan")), torch.full_like(qkv, float("nan"))
    delta_ref = torch.empty(M, H, device="cuda")
    ops.cos_attn_bwd_qknorm(qkv, att, dh, lse, sc, d_ref, delta_ref, N, T, H, hd)
    delta2 = torch.full((M, H), float("nan"), device="cuda")
    ops.gemm_bf16(dy, wt, dh, epilogue=_lib.EPI_STORE_DELTA, resid=att, aux=delta2)
    assert rel_l2(delta2, delta_ref) < 1e-5
    ops.cos_attn_bwd_qknorm(qkv, None, dh, lse, sc, d_new, delta2, N, T, H, hd)
    assert rel_l2(d_new.float(), d_ref.float()) < 2e-3
    with pytest.raises(RuntimeError):  # delta can only be handed over on the fused tokens == 256 path
        ops.cos_attn_bwd_qknorm(qkv[: 2 * 64], None, dh[: 2 * 64], lse[: 2 * 64], sc[: 2 * 64], d_new[: 2 * 64], delta2[: 2 * 64], 2, 64, H, hd)


@pytest.mark.parametrize("N,T,H", [(2, 256, 3), (1, 1024, 2), (3, 512, 16), (40, 256, 4)])
def test_cos_attn_bf16_head_dim_72(N, T, H):
    """DiT-XL's head_dim 72 on the tcgen05 forward kernel: two 64-channel panels per operand tile, the second one zero-filled past
    channel 72 by the TMA unit (3-D tensor map), S over 80 channels, PV with 80 output columns, row sums from the softmax warps"""
    from mapdit_b200 import ops
    hd, D = 72, H * 72
    qkv = rnd(N * T, 3 * D, seed=19)
    ops.qk_normalize(qkv, D, hd)
    qkv16 = qkv.bfloat16()
    q, k, v = qkv16.float().view(N, T, 3, H, hd).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q.double(), k.double(), v.double(), scale=1 / math.sqrt(hd)).transpose(1, 2).reshape(N * T, D)
    ref_lse = torch.logsumexp(q.double() @ k.double().transpose(-1, -2) / math.sqrt(hd), dim=-1).permute(0, 2, 1).reshape(N * T, H)
    o = torch.full((N * T, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    lse = torch.full((N * T, H), float("nan"), device="cuda")
    ops.cos_attn(qkv16, o, N, T, H, hd, lse=lse)
    torch.cuda.synchronize()
    e = rel_l2(o.float(), ref)
    print(f"head_dim 72 tcgen05 attention N={N} T={T} H={H}: rel-L2 {e:.2e}, lse max abs err {float((lse.double() - ref_lse).abs().max()):.2e}")
    assert e < 6e-3
    assert float((lse.double() - ref_lse).abs().max()) < 2e-2
