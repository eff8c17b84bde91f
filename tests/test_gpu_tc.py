"""-m gpu: tcgen05 tensor-core kernels (GEMM + fused epilogues, QKV projection, attention) through the C ABI,
against torch fp32 references computed from the same bf16-rounded operands (fp32 accumulation on both sides)."""
import pytest
import torch
import torch.nn.functional as F

from boxdreamer_b200 import _lib
from gpu_util import gemm, report, sp

pytestmark = pytest.mark.gpu
TC = _lib.PRECISION_BF16


def _operands(M, N, K, seed):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16)
    b = torch.randn(N, generator=g)
    ref = (A.double() @ W.double().t() + b.double()).float()
    return A.cuda(), W.cuda(), b.cuda(), ref


@pytest.fixture
def gemm_pair():
    """(kept as a fixture name: every GEMM runs through the CTA-pair cta_group::2 kernel, gemm_tc2.cu)"""
    return 1


@pytest.mark.parametrize("M,N,K", [
    (128, 256, 64),      # one tile, one k-block
    (128, 256, 256),     # one tile, 4 k-blocks (ring wrap-free)
    (256, 512, 768),     # 2x2 tiles, 12 k-blocks (ring wraps, both accumulator stages)
    (1000, 768, 640),    # M tail, patch-embed K
    (512, 1568, 768),    # N tail (heat-map head)
    (384, 768, 1568),    # K tail (bbox_emb)
    (20000, 768, 3072),  # more tiles than SMs (persistent loop), fc2 shape
])
def test_gemm_tc_f32(lib, gemm_pair, M, N, K):
    A, W, b, ref = _operands(M, N, K, M + N + K)
    out = gemm(A, W, b, M, N, K, _lib.EPI_F32, TC)
    ok, msg = report(f"gemm_tc_f32[{M},{N},{K}]", out, ref, tol_rel=1e-5)
    assert ok, msg


@pytest.mark.parametrize("M,N,K", [(256, 768, 768), (1000, 3072, 768)])
def test_gemm_tc_epilogues(lib, gemm_pair, M, N, K):
    A, W, b, ref = _operands(M, N, K, 3 * M + N)
    out = gemm(A, W, b, M, N, K, _lib.EPI_ACT, TC)
    ok, msg = report("gemm_tc_act(bf16)", out, ref, tol_rel=6e-3)
    assert ok, msg
    out = gemm(A, W, b, M, N, K, _lib.EPI_GELU, TC)
    ok, msg = report("gemm_tc_gelu(bf16)", out, F.gelu(ref), tol_rel=6e-3)
    assert ok, msg
    g = torch.Generator().manual_seed(9)
    res0 = torch.randn(M, N, generator=g).cuda()
    gam = torch.randn(N, generator=g).cuda()
    out = gemm(A, W, b, M, N, K, _lib.EPI_RESID, TC, gamma=gam, out=res0.clone())
    ok, msg = report("gemm_tc_resid", out, res0.cpu() + gam.cpu() * ref, tol_rel=1e-5)
    assert ok, msg
    out = gemm(A, W, b, M, N, K, _lib.EPI_RESID, TC, gamma=None, out=res0.clone())
    ok, msg = report("gemm_tc_resid_nogamma", out, res0.cpu() + ref, tol_rel=1e-5)
    assert ok, msg


def _qkv_case(L, seq, heads, hd, norm, seed):
    from oracle import boxdreamer_oracle as O
    d = heads * hd
    seq_pad = (seq + 127) // 128 * 128
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(L * seq, d, generator=g).to(torch.bfloat16)
    W = (torch.randn(3 * d, d, generator=g) * 0.04).to(torch.bfloat16)
    b = torch.randn(3 * d, generator=g) * 0.05
    qw = 1 + 0.1 * torch.randn(hd, generator=g)
    kw = 1 + 0.1 * torch.randn(hd, generator=g)
    qkv = F.linear(x.float(), W.float(), b).view(L, seq, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    if norm:
        q, k = O.rms_norm(q, qw), O.rms_norm(k, kw)
    return x, W, b, qw, kw, q, k, v, seq_pad


@pytest.mark.parametrize("L,seq,heads,hd,norm", [(2, 512, 8, 96, True), (3, 261, 12, 64, False), (1, 1536, 8, 96, True), (2, 500, 8, 96, True),
                                                 (2, 512, 12, 64, False), (3, 256, 12, 64, True), (1, 261, 12, 64, False)])
def test_qkv_project_tc(lib, gemm_pair, L, seq, heads, hd, norm):
    x, W, b, qw, kw, q, k, v, seq_pad = _qkv_case(L, seq, heads, hd, norm, 17)
    Q = torch.zeros(L * heads, seq_pad, hd, device="cuda", dtype=torch.bfloat16)
    K = torch.zeros_like(Q)
    Vt = torch.zeros(L * heads, hd, seq_pad, device="cuda", dtype=torch.bfloat16)
    xc, Wc, bc = x.cuda(), W.cuda(), b.cuda()
    qwc, kwc = (qw.cuda(), kw.cuda()) if norm else (None, None)
    _lib.check(lib.bd_qkv_project(_lib.ptr(xc), _lib.ptr(Wc), _lib.ptr(bc), _lib.ptr(qwc), _lib.ptr(kwc), _lib.ptr(Q), _lib.ptr(K),
                                  _lib.ptr(Vt), None, L, seq, seq_pad, heads, hd, TC, sp()))
    torch.cuda.synchronize()
    for name, got, ref in (("Q", Q.view(L, heads, seq_pad, hd)[:, :, :seq], q), ("K", K.view(L, heads, seq_pad, hd)[:, :, :seq], k),
                           ("Vt", Vt.view(L, heads, hd, seq_pad)[:, :, :, :seq].transpose(2, 3), v)):
        ok, msg = report(f"qkv_tc.{name}", got, ref, tol_rel=6e-3)
        assert ok, msg
    # padding must stay untouched (zero)
    assert float(Q.view(L, heads, seq_pad, hd)[:, :, seq:].abs().max() if seq_pad > seq else 0) == 0.0


ATT_CASES = [(1, 128, 8, 96), (2, 512, 8, 96), (3, 261, 12, 64), (1, 1536, 8, 96), (2, 700, 8, 96)]


@pytest.mark.parametrize("L,seq,heads,hd", ATT_CASES + [(5, 261, 12, 64), (3, 100, 8, 96), (1, 2000, 8, 96), (40, 261, 12, 64)])
def test_attention_tc_pingpong(lib, L, seq, heads, hd):
    _attention_case(lib, 2, L, seq, heads, hd)


def _attention_case(lib, variant, L, seq, heads, hd):
    seq_pad = (seq + 127) // 128 * 128
    g = torch.Generator().manual_seed(seq + hd + variant)
    q = torch.randn(L, heads, seq, hd, generator=g).to(torch.bfloat16)
    k = torch.randn(L, heads, seq, hd, generator=g).to(torch.bfloat16)
    v = torch.randn(L, heads, seq, hd, generator=g).to(torch.bfloat16)
    Q = torch.zeros(L * heads, seq_pad, hd, dtype=torch.bfloat16)
    K = torch.zeros_like(Q)
    Vt = torch.zeros(L * heads, hd, seq_pad, dtype=torch.bfloat16)
    Q.view(L, heads, seq_pad, hd)[:, :, :seq] = q
    K.view(L, heads, seq_pad, hd)[:, :, :seq] = k
    Vt.view(L, heads, hd, seq_pad)[:, :, :, :seq] = v.transpose(2, 3)
    Q, K, Vt = Q.cuda(), K.cuda(), Vt.cuda()
    Oo = torch.zeros(L * seq, heads * hd, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.bd_attention(_lib.ptr(Q), _lib.ptr(K), _lib.ptr(Vt), _lib.ptr(Oo), L, heads, hd, seq, seq_pad, hd ** -0.5, TC,
                                variant, sp()))
    torch.cuda.synchronize()
    ref = F.scaled_dot_product_attention(q.float(), k.float(), v.float(), scale=hd ** -0.5).transpose(1, 2).reshape(L * seq, heads * hd)
    # P is rounded to bf16 before the PV product and the output is bf16: 1e-2 of the output range
    ok, msg = report(f"attention_tc[v{variant}]", Oo, ref, tol_rel=1.5e-2)
    assert ok, msg
