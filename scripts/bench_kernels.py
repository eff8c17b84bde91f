"""Kernel-level timings through the C ABI (CUDA events on the launching stream, L2-exceeding operands):
attention variants at the decoder / DINOv2 shapes of BASELINE config 2 and the five GEMM shapes."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boxdreamer_b200 import _lib
lib = _lib.load()
TC = _lib.PRECISION_BF16


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def attention(L, heads, hd, seq, variants=(2,)):
    seq_pad = (seq + 127) // 128 * 128
    Q = torch.randn(L * heads, seq_pad, hd, device="cuda").to(torch.bfloat16)
    K = torch.randn_like(Q)
    Vt = torch.randn(L * heads, hd, seq_pad, device="cuda").to(torch.bfloat16)
    O = torch.empty(L * seq, heads * hd, device="cuda", dtype=torch.bfloat16)
    flops = 4.0 * seq * seq * hd * heads * L
    out = {}
    ref = None
    for v in variants:
        def run():
            _lib.check(lib.bd_attention(_lib.ptr(Q), _lib.ptr(K), _lib.ptr(Vt), _lib.ptr(O), L, heads, hd, seq, seq_pad, hd ** -0.5, TC, v,
                                        _lib.stream_ptr()))
        try:
            ms = timeit(run)
            o = O.float().clone()
            if ref is None:
                ref = o
            out[f"v{v}"] = {"ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1), "maxdiff_vs_first": float((o - ref).abs().max())}
        except Exception as e:  # keep going: a failing variant must not hide the others
            out[f"v{v}"] = {"error": str(e)[:200]}
            break
    return out


def gemm(M, N, K, epi):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    b = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if epi in (0, 2) else torch.bfloat16)
    def run():
        _lib.check(lib.bd_gemm(_lib.ptr(A), _lib.ptr(W), _lib.ptr(b), None, _lib.ptr(out), M, N, K, epi, TC, _lib.stream_ptr()))
    ms = timeit(run)
    return {"ms": round(ms, 4), "tflops": round(2.0 * M * N * K / ms / 1e9, 1)}


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "attn":   # attention only, current variant
        print(json.dumps({
            "attn_decoder_L64_h8_d96_N1536": attention(64, 8, 96, 1536, (2,)),
            "attn_dino_L384_h12_d64_N261": attention(384, 12, 64, 261, (2,)),
            "attn_long_L4_h8_d96_N9792": attention(4, 8, 96, 9792, (2,)),
            "attn_336_L34_h12_d64_N581": attention(34, 12, 64, 581, (2,)),
        }, indent=1))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "gemm":
        for name, r in (("proj_98304x768x768_resid", gemm(98304, 768, 768, 2)), ("fc1_98304x3072x768_gelu", gemm(98304, 3072, 768, 1)),
                        ("fc2_98304x768x3072_resid", gemm(98304, 768, 3072, 2)), ("f32_98304x768x768", gemm(98304, 768, 768, 0)),
                        ("act_98304x3072x3072", gemm(98304, 3072, 3072, 4)), ("dino_fc2_100224x768x3072_resid", gemm(100224, 768, 3072, 2))):
            print(f" {name:36s} {r['ms']:8.4f} ms  {r['tflops']:7.1f} TFLOP/s")
        sys.exit(0)
    res = {
        "attn_decoder_L64_h8_d96_N1536": attention(64, 8, 96, 1536, (2,)),
        "attn_dino_L384_h12_d64_N261": attention(384, 12, 64, 261, (2,)),
        "attn_long_L4_h8_d96_N9792": attention(4, 8, 96, 9792, (2,)),
        "gemm_proj_98304x768x768_resid": gemm(98304, 768, 768, 2),
        "gemm_fc1_98304x3072x768_gelu": gemm(98304, 3072, 768, 1),
        "gemm_fc2_98304x768x3072_resid": gemm(98304, 768, 3072, 2),
        "gemm_f32_98304x768x768": gemm(98304, 768, 768, 0),
    }
    print(json.dumps(res, indent=1))
