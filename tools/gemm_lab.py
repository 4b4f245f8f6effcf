"""Where does the pair GEMM's main loop lose its cycles?  Runs the CTA-pair kernel with the measurement knobs of
GemmArgs::lab (no operand refill = issue-loop ceiling; all loads from tile (0,0) = L2-hit-only supply; L2 prefetch
distance) and reads the MMA issuer's in-kernel wait counters -> gpurun_out/gemm_lab.jsonl."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bert_pytorch_b200.ops import api as K  # noqa: E402
from bert_pytorch_b200.ops import _loader  # noqa: E402

C = _loader.load_extension()


def timeit(fn, iters=12, warm=4):
    flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.fill_(0.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    M = 12288
    shapes = [("big", K.NT, 8192, 8192, 8192), ("qkv_fwd", K.NT, M, 3072, 1024), ("ffn2_fwd", K.NT, M, 1024, 4096),
              ("ffn1_dgrad", K.NN, M, 1024, 4096), ("ffn2_dgrad", K.NN, M, 4096, 1024), ("ffn1_wgrad", K.TN, 4096, 1024, M),
              ("ffn2_wgrad", K.TN, 1024, 4096, M)]
    variants = [("base", 0), ("norefill", 1), ("noepi", 4), ("ldonly", 8), ("norefill_noepi", 5), ("norefill_ldonly", 9),
                ("stag1", 16 | (1 << 16)), ("stag4", 16 | (4 << 16)), ("stag13", 16 | (13 << 16)), ("box3d", 32),
                ("box3d_stag4", 48 | (4 << 16)), ("nomma", 64 | 4), ("nomma_tile00", 64 | 4 | 2), ("nomma_box3d", 64 | 4 | 32),
                ("direct_epi", 128), ("norefill_direct_epi", 129), ("oldroles", 256), ("oldroles_noepi", 256 | 4),
                ("nomma_37", 64 | 4 | (37 << 24)), ("nomma_18", 64 | 4 | (18 << 24)), ("noepi_37", 4 | (37 << 24)),
                ("base_37", 37 << 24)]
    if "--variants" in sys.argv:
        want = sys.argv[sys.argv.index("--variants") + 1].split(",")
        variants = [v for v in variants if v[0] in want]
    stats = torch.zeros(74 * 12, dtype=torch.int64, device="cuda")
    # the knobs and clocks exist in lab builds only (B200_NVCC_EXTRA=-DB200_GEMM_LAB python -m bert_pytorch_b200.ops.build)
    probe_a, probe_b = torch.randn(512, 256, device="cuda").bfloat16(), torch.randn(512, 256, device="cuda").bfloat16()
    C.gemm_lab(0, stats)
    K.gemm(probe_a, probe_b, block_n=512)
    torch.cuda.synchronize()
    C.gemm_lab(0, None)
    if int(stats.abs().sum()) == 0:
        sys.exit("gemm_lab: this build has no measurement knobs -- rebuild with B200_NVCC_EXTRA=-DB200_GEMM_LAB "
                 "(python -m bert_pytorch_b200.ops.build), ideally in a scratch worktree: the knobs cost the production "
                 "kernels up to 19 %")
    out_path = os.path.join("gpurun_out", "gemm_lab.jsonl")
    os.makedirs("gpurun_out", exist_ok=True)
    with open(out_path, "w") as f:
        for name, layout, m, n, k in shapes:
            if layout == K.NT:
                a, b = torch.randn(m, k, device="cuda").bfloat16(), torch.randn(n, k, device="cuda").bfloat16()
                ref = lambda: a @ b.t()
            elif layout == K.NN:
                a, b = torch.randn(m, k, device="cuda").bfloat16(), torch.randn(k, n, device="cuda").bfloat16()
                ref = lambda: a @ b
            else:
                a, b = torch.randn(k, m, device="cuda").bfloat16(), torch.randn(k, n, device="cuda").bfloat16()
                ref = lambda: a.t() @ b
            flops = 2.0 * m * n * k
            rec = {"name": name, "M": m, "N": n, "K": k, "cublas_tflops": round(flops / timeit(ref) / 1e9, 1)}
            if layout == K.TN:
                o = torch.zeros(m, n, device="cuda")
                run = lambda: K.gemm(a, b, layout=layout, epi=K.EPI_ACCUM_F32, out=o, block_n=512, k_splits=1)
            else:
                o = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
                run = lambda: K.gemm(a, b, layout=layout, out=o, block_n=512)
            for vname, flags in variants:
                C.gemm_lab(flags, None)
                t = timeit(run)
                stats.zero_()
                C.gemm_lab(flags, stats)
                run()
                torch.cuda.synchronize()
                st = stats[:74 * 4].view(74, 4).double()
                live = st[:, 3] > 0
                tot, wf, wt, kb = (st[live, i] for i in range(4))
                rec[vname] = {"tflops": round(flops / t / 1e9, 1), "ms": round(t, 4),
                              "cyc_per_kblock": round(float((tot / kb).mean()), 1),
                              "mhz": round(float(tot.max()) / (t * 1e3), 0),
                              "wait_tmem_frac": round(float((wt / tot).mean()), 3)}
            C.gemm_lab(0, None)
            print(json.dumps(rec), flush=True)
            f.write(json.dumps(rec) + "\n")
        if "--wgrad" in sys.argv:               # the weight-gradient calls exactly as the engine issues them (split-K etc.)
            for name, n_out, k_out in (("ffn1_wgrad", 4096, 1024), ("ffn2_wgrad", 1024, 4096), ("qkv_wgrad", 3072, 1024),
                                       ("attn_out_wgrad", 1024, 1024)):
                dy = torch.randn(M, n_out, device="cuda").bfloat16()
                x = torch.randn(M, k_out, device="cuda").bfloat16()
                grad = torch.zeros(n_out, k_out, device="cuda")
                rec = {"name": name + "_engine", "splits": K.wgrad_splits(n_out, k_out, M, 512)}
                for vname, flags in (("box3d", 0), ("box2d", 32), ("box3d_oldroles", 256), ("box2d_oldroles", 32 | 256)):
                    C.gemm_lab(flags, None)
                    rec[vname + "_ms"] = round(timeit(lambda: K.wgrad_accumulate(dy, x, grad)), 4)
                C.gemm_lab(0, None)
                print(json.dumps(rec), flush=True)
                f.write(json.dumps(rec) + "\n")
        if "--epi" in sys.argv:                 # the fused epilogues of the engine, in issuer cycles per k-block
            def measure(run, flags=0):
                C.gemm_lab(flags, None)
                t = timeit(run)
                stats.zero_()
                C.gemm_lab(flags, stats)
                run()
                torch.cuda.synchronize()
                C.gemm_lab(0, None)
                st = stats[:74 * 4].view(74, 4).double()
                live = st[:, 3] > 0
                ph = stats[74 * 4:].view(74, 8).double()[live]
                ntile = (st[live, 3] / max(k // 64, 1)).clamp(min=1).unsqueeze(1)
                # cycles per tile of the box-pipelined epilogue (first thread of half 0, both boxes summed): wait tmem_full |
                # wait residual / box free | math | half barrier | store (+colsum) | wait store read + next residual | - | drain
                return {"ms": round(t, 4), "cyc_per_kblock": round(float((st[live, 0] / st[live, 3]).mean()), 1),
                        "wait_tmem_frac": round(float((st[live, 2] / st[live, 0]).mean()), 3),
                        "epi_phase_cyc_per_tile": [int(x) for x in (ph / ntile).mean(0).tolist()]}
            m = M
            for name, layout, n, k in (("ffn1_fwd", K.NT, 4096, 1024), ("attn_out_fwd", K.NT, 1024, 1024),
                                       ("ffn2_fwd", K.NT, 1024, 4096), ("ffn2_dgrad", K.NN, 4096, 1024)):
                a = torch.randn(m, k, device="cuda").bfloat16()
                b = (torch.randn(n, k, device="cuda") if layout == K.NT else torch.randn(k, n, device="cuda")).bfloat16()
                o = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
                aux = torch.empty_like(o)
                bias = torch.randn(n, device="cuda").bfloat16()
                res = torch.randn(m, n, device="cuda").bfloat16()
                cs = torch.zeros(n, device="cuda")
                rec = {"name": name + "_epi", "M": m, "N": n, "K": k}
                rec["plain"] = measure(lambda: K.gemm(a, b, layout=layout, out=o, block_n=512))
                rec["plain_oldroles"] = measure(lambda: K.gemm(a, b, layout=layout, out=o, block_n=512), 256)
                rec["bias"] = measure(lambda: K.gemm(a, b, layout=layout, out=o, epi=K.EPI_BIAS, bias=bias, block_n=512))
                if name == "ffn1_fwd":
                    rec["bias_gelu_dg"] = measure(lambda: K.gemm(a, b, layout=layout, out=o, epi=K.EPI_BIAS_GELU_DG, bias=bias,
                                                                 aux_out=aux, block_n=512))
                    rec["bias_gelu_dg_oldroles"] = measure(lambda: K.gemm(a, b, layout=layout, out=o, epi=K.EPI_BIAS_GELU_DG,
                                                                          bias=bias, aux_out=aux, block_n=512), 256)
                elif name == "ffn2_dgrad":
                    rec["mul"] = measure(lambda: K.gemm(a, b, layout=layout, out=o, epi=K.EPI_MUL, res=res, block_n=512))
                    rec["mul_colsum"] = measure(lambda: K.gemm(a, b, layout=layout, out=o, epi=K.EPI_MUL, res=res, colsum=cs,
                                                               block_n=512))
                else:
                    rec["bias_drop_res"] = measure(lambda: K.gemm(a, b, layout=layout, out=o, epi=K.EPI_BIAS_DROP_RES, bias=bias,
                                                                  res=res, p_drop=0.1, seed=1234, block_n=512))
                    bits = K.dropout_mask(m, n, 0.1, 1234, 0, "cuda")
                    rec["bias_drop_res_maskin"] = measure(lambda: K.gemm(a, b, layout=layout, out=o, epi=K.EPI_BIAS_DROP_RES,
                                                                         bias=bias, res=res, p_drop=0.1, seed=1234, block_n=512,
                                                                         mask_in=bits))
                    rec["add_res"] = measure(lambda: K.gemm(a, b, layout=layout, out=o, epi=K.EPI_ADD, res=res, block_n=512))
                print(json.dumps(rec), flush=True)
                f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
