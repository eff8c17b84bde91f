"""The bar on the same box: the reference's attention call -- `flash_attn_func` (FA2,
blocks.py:259-272) and its fallback `F.scaled_dot_product_attention` (blocks.py:273-285) -- on the decoder / long-sequence
/ DINOv2 shapes, beside this repo's attn_tc2 kernel through the C ABI.  CUDA events, operands larger than L2 where the shape
allows, clocks sampled during the run.  FLOPs = 4 * N^2 * hd * heads * L (QK^T and PV only).

    python scripts/bench_attn_bar.py > gpurun_out/attn_bar.json
"""
import json, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boxdreamer_b200 import _lib
from bench import ClockSampler

lib = _lib.load()


def timeit(fn, iters=20, warm=5):
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


def one_shape(L, heads, hd, seq):
    flops = 4.0 * seq * seq * hd * heads * L
    out = {"shape": {"L": L, "heads": heads, "head_dim": hd, "seq": seq}, "gflop_per_call": flops / 1e9}
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(L, seq, heads, hd, device="cuda", generator=g).to(torch.bfloat16)   # flash layout [B, N, H, hd] (blocks.py:263-272)
    k = torch.randn(L, seq, heads, hd, device="cuda", generator=g).to(torch.bfloat16)
    v = torch.randn(L, seq, heads, hd, device="cuda", generator=g).to(torch.bfloat16)
    ref = None
    try:
        from flash_attn import flash_attn_func
        ms = timeit(lambda: flash_attn_func(q, k, v, dropout_p=0.0, softmax_scale=hd ** -0.5, causal=False))
        ref = flash_attn_func(q, k, v, dropout_p=0.0, softmax_scale=hd ** -0.5, causal=False).float()
        out["flash_attn_func"] = {"ms": ms, "tflops": flops / ms / 1e9}
        import flash_attn
        out["flash_attn_func"]["version"] = getattr(flash_attn, "__version__", "?")
    except Exception as exc:
        out["flash_attn_func"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:160]}
    # SDPA takes [B, H, N, hd] (blocks.py:279-285 passes the permuted q, k, v)
    qs, ks, vs = (t.permute(0, 2, 1, 3).contiguous() for t in (q, k, v))
    try:
        ms = timeit(lambda: F.scaled_dot_product_attention(qs, ks, vs, dropout_p=0.0, scale=hd ** -0.5))
        o = F.scaled_dot_product_attention(qs, ks, vs, dropout_p=0.0, scale=hd ** -0.5)
        if ref is None:
            ref = o.permute(0, 2, 1, 3).float()
        out["sdpa_default_backend"] = {"ms": ms, "tflops": flops / ms / 1e9}
    except Exception as exc:
        out["sdpa_default_backend"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:160]}
    for name, backend in (("sdpa_flash", "FLASH_ATTENTION"), ("sdpa_cudnn", "CUDNN_ATTENTION"), ("sdpa_efficient", "EFFICIENT_ATTENTION")):
        try:
            from torch.nn.attention import SDPBackend, sdpa_kernel
            with sdpa_kernel(getattr(SDPBackend, backend)):
                ms = timeit(lambda: F.scaled_dot_product_attention(qs, ks, vs, dropout_p=0.0, scale=hd ** -0.5))
            out[name] = {"ms": ms, "tflops": flops / ms / 1e9}
        except Exception as exc:
            out[name] = {"unavailable": f"{type(exc).__name__}: {exc}"[:160]}
    # ours: Q, K [BH, seq_pad, hd], V^T [BH, hd, seq_pad] (the QKV epilogue's layouts), O token-major
    seq_pad = (seq + 127) // 128 * 128
    Q = torch.zeros(L * heads, seq_pad, hd, device="cuda", dtype=torch.bfloat16)
    K = torch.zeros_like(Q)
    Vt = torch.zeros(L * heads, hd, seq_pad, device="cuda", dtype=torch.bfloat16)
    Q[:, :seq] = qs.reshape(L * heads, seq, hd)
    K[:, :seq] = ks.reshape(L * heads, seq, hd)
    Vt[:, :, :seq] = vs.reshape(L * heads, seq, hd).transpose(1, 2)
    O = torch.empty(L * seq, heads * hd, device="cuda", dtype=torch.bfloat16)

    def ours():
        _lib.check(lib.bd_attention(_lib.ptr(Q), _lib.ptr(K), _lib.ptr(Vt), _lib.ptr(O), L, heads, hd, seq, seq_pad, hd ** -0.5,
                                    _lib.PRECISION_BF16, 2, _lib.stream_ptr()))
    ms = timeit(ours)
    out["attn_tc2 (this repo)"] = {"ms": ms, "tflops": flops / ms / 1e9}
    if ref is not None:
        out["attn_tc2 (this repo)"]["max_abs_diff_vs_reference_kernel"] = float((O.view(L, seq, heads, hd).float() - ref).abs().max())
    base = out.get("flash_attn_func", {}).get("ms") or out.get("sdpa_default_backend", {}).get("ms")
    if base:
        out["speedup_vs_reference_kernel"] = base / ms
    return out


if __name__ == "__main__":
    sampler = ClockSampler(0)
    sampler.start()
    res = {"decoder_config2": one_shape(64, 8, 96, 1536), "decoder_config4_long": one_shape(4, 8, 96, 9792),
           "dino_224": one_shape(384, 12, 64, 261), "dino_336": one_shape(68, 12, 64, 581)}
    res["clocks"] = sampler.stop()
    res["torch"] = torch.__version__
    res["gpu"] = torch.cuda.get_device_name(0)
    print(json.dumps(res, indent=1))
