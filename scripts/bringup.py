"""GPU bring-up: checks each stage of the CUDA path against torch on the same device.
Usage (on the B200 box): python scripts/bringup.py [stage ...]   stages: gemm stats grad critic"""
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mi_b200  # noqa: E402
from mi_b200 import ops, _lib  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
print("device:", torch.cuda.get_device_name(0), "abi", lib.mi_abi_version(), "check", lib.mi_device_check(), flush=True)


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def stage_gemm(cg):
    lib.mi_set_cta_group(cg)
    ok = True
    for (M, N, K) in [(128, 256, 64), (256, 256, 128), (300, 520, 200), (1024, 768, 1024), (4096, 1024, 4096)]:
        g = torch.Generator(device="cpu").manual_seed(M + N + K)
        A = torch.randn(M, K, generator=g).to(dev).bfloat16()
        B = torch.randn(N, K, generator=g).to(dev).bfloat16()
        C = ops.gemm(A, B)
        torch.cuda.synchronize()
        ref = A.float() @ B.float().t()
        e = relerr(C, ref)
        sub = torch.randn(M, N, generator=g).to(dev).bfloat16()
        C16 = ops.gemm(A, B, alpha=0.5, gamma=2.0, sub=sub, out_dtype=torch.bfloat16)
        torch.cuda.synchronize()
        e16 = relerr(C16.float(), 0.5 * (ref - 2.0 * sub.float()))
        print(f"  gemm cg={cg} {M}x{N}x{K}: relerr {e:.2e}  (alpha/gamma/sub, bf16 out: {e16:.2e})", flush=True)
        ok &= e < 1e-5 and e16 < 1e-2
    return ok


def ref_stats(Q, K, sid_q, sid_k, q_offset, scale):
    S = (Q.double() @ K.double().t()) * scale
    M = sid_q[:, None] != sid_k[None, :]
    idx = torch.arange(Q.shape[0], device=Q.device)
    diag = S[idx, idx + q_offset]
    Sm = torch.where(M, S, torch.full_like(S, float("-inf")))
    lse_neg = torch.logsumexp(Sm, 1)
    n_neg = M.sum(1).double()
    lse_all = torch.logaddexp(lse_neg, diag)
    return S, M, lse_neg, n_neg, diag, lse_all


def stage_stats(cg):
    lib.mi_set_cta_group(cg)
    ok = True
    for (Bq, Bk, D, off) in [(32, 32, 64, 0), (200, 200, 72, 0), (128, 512, 256, 256), (1000, 3000, 768, 1500), (4096, 4096, 1024, 0)]:
        g = torch.Generator().manual_seed(Bq + Bk)
        Q = (torch.randn(Bq, D, generator=g) / D ** 0.25).to(dev).bfloat16()
        K = (torch.randn(Bk, D, generator=g) / D ** 0.25).to(dev).bfloat16()
        sid_k = torch.arange(Bk, dtype=torch.int32)
        sid_k[1::7] = sid_k[0::7][: len(sid_k[1::7])]
        sid_k = sid_k.to(dev)
        sid_q = sid_k[off:off + Bq].clone()
        rows, scal = ops.score_stats(Q, K, sid_q, sid_k, off, 0.7)
        torch.cuda.synchronize()
        S, M, lse_neg, n_neg, diag, lse_all = ref_stats(Q, K, sid_q, sid_k, off, 0.7)
        e1 = float((rows[:, 0].double() - lse_neg).abs().max())
        e2 = float((rows[:, 1].double() - n_neg).abs().max())
        e3 = float((rows[:, 2].double() - diag).abs().max())
        e4 = float((rows[:, 3].double() - lse_all).abs().max())
        gl = torch.logsumexp(lse_neg, 0)
        e5 = abs(float(scal[0] + torch.log(scal[1])) - float(gl))
        e6 = abs(float(scal[2]) - float(n_neg.sum())) + abs(float(scal[3]) - float(diag.sum()))
        print(f"  stats cg={cg} Bq={Bq} Bk={Bk} D={D}: lse_neg {e1:.2e} n_neg {e2:.0f} diag {e3:.2e} lse_all {e4:.2e} glse {e5:.2e} sums {e6:.2e}", flush=True)
        ok &= e1 < 2e-4 and e2 == 0 and e3 < 2e-4 and e4 < 2e-4 and e5 < 2e-4
    return ok


def stage_grad(cg):
    lib.mi_set_cta_group(cg)
    ok = True
    for (Bq, Bk, D, off, strict) in [(64, 64, 64, 0, "fast"), (200, 200, 72, 0, "fast"), (300, 900, 256, 300, "strict"),
                                     (2048, 2048, 768, 0, "fast"), (2048, 2048, 768, 0, "strict"), (1000, 5000, 1024, 2000, "fast")]:
        g = torch.Generator().manual_seed(Bq + Bk + 1)
        Q = (torch.randn(Bq, D, generator=g) / D ** 0.25).to(dev).bfloat16()
        K = (torch.randn(Bk, D, generator=g) / D ** 0.25).to(dev).bfloat16()
        sid_k = torch.arange(Bk, dtype=torch.int32)
        sid_k[1::5] = sid_k[0::5][: len(sid_k[1::5])]
        sid_k = sid_k.to(dev)
        sid_q = sid_k[off:off + Bq].clone()
        S, M, lse_neg, n_neg, diag, lse_all = ref_stats(Q, K, sid_q, sid_k, off, 0.7)
        idx = torch.arange(Bq, device=dev)
        R = M.clone()
        R[idx, idx + off] = True
        refq = lse_all.float()
        col_lse = torch.logsumexp(torch.where(R, S, torch.full_like(S, float("-inf"))), 0)
        refk = col_lse.float()
        wq, wk = 0.5 / Bk, 0.25 / Bk
        G = torch.where(R, wq * torch.exp(S - refq.double()[:, None]) + wk * torch.exp(S - refk.double()[None, :]), torch.zeros_like(S))
        gam = 1.0 / Bk
        refo = 0.7 * (G @ K.double() - gam * K[off:off + Bq].double())
        refkk = G.t() @ Q.double()
        refkk[off:off + Bq] -= gam * Q.double()
        refkk = 0.7 * refkk
        o32, o16, okk = ops.score_grad(Q, K, sid_q, sid_k, off, 0.7, refq, wq, refk, wk, True, strict, 0.7, gam,
                                       want_f32=True, want_bf16=True, out_split=(strict == "strict"), want_k=True)
        torch.cuda.synchronize()
        e = relerr(o32, refo)
        e16 = relerr(o16.float(), refo)
        ek = relerr(okk, refkk)
        # DV-style: negatives only, scalar reference
        glse = torch.logsumexp(lse_neg, 0)
        Gd = torch.where(M, torch.exp(S - glse), torch.zeros_like(S))
        refd = Gd @ K.double() - gam * K[off:off + Bq].double()
        od, _, _ = ops.score_grad(Q, K, sid_q, sid_k, off, 0.7, torch.full((Bq,), float(glse), device=dev), 1.0, None, 0.0,
                                  False, strict, 1.0, gam)
        torch.cuda.synchronize()
        ed = relerr(od, refd)
        print(f"  grad cg={cg} Bq={Bq} Bk={Bk} D={D} {strict}: Oq {e:.2e} (bf16/split out {e16:.2e})  Ok(MN-major) {ek:.2e}  dv-form {ed:.2e}", flush=True)
        tol = 3e-3 if strict == "fast" else 2e-4
        ok &= e < tol and ed < tol and ek < tol
    return ok


def stage_critic(cg):
    lib.mi_set_cta_group(cg)
    from oracle import matrix_oracle as mo
    ok = True
    for (B, D, critic, est, prec) in [(32, 768, "dot", "dv", "strict"), (96, 64, "bilinear", "dv", "strict"),
                                      (200, 128, "bilinear", "infonce", "fast"), (256, 256, "bilinear", "infonce_row", "strict"),
                                      (300, 72, "dot", "infonce_sym", "strict"), (1024, 768, "bilinear", "infonce_sym", "fast")]:
        X, Y, sid, W = mo.synthetic_embeddings(B, D, seed=B, dup_frac=0.05, bilinear=(critic == "bilinear"))
        Xb, Yb = X.bfloat16(), Y.bfloat16()
        Wb = None if W is None else W.bfloat16()
        inv_tau = 1.0 / math.sqrt(D) if critic == "dot" else 1.0
        ref = mo.critic_loss(Xb.float(), Yb.float(), sid, None if Wb is None else Wb.float(), inv_tau, est)
        out, dX, dY, dW = ops.critic_loss_fwd_bwd(Xb.to(dev), Yb.to(dev), None if Wb is None else Wb.to(dev), sid.to(dev), est, prec, inv_tau, True)
        torch.cuda.synchronize()
        el = abs(float(out[0]) - float(ref["loss"])) / max(abs(float(ref["loss"])), 1e-12)
        ex, ey = relerr(dX.cpu(), ref["dX"]), relerr(dY.cpu(), ref["dY"])
        ew = relerr(dW.cpu(), ref["dW"]) if dW is not None else 0.0
        print(f"  critic cg={cg} B={B} D={D} {critic} {est} {prec}: loss {float(out[0]):.6f} vs {float(ref['loss']):.6f} rel {el:.2e}  dX {ex:.2e} dY {ey:.2e} dW {ew:.2e}  n_neg {float(out[3]):.0f}/{float(ref['n_neg']):.0f}", flush=True)
        tol = (1e-4, 1e-3) if prec == 'strict' else (2e-3, 1e-2)
        ok &= el / max(1.0, 1.0 / max(abs(float(ref['loss'])), 1e-12)) < tol[0] and max(ex, ey, ew) < tol[1] and float(out[3]) == float(ref["n_neg"])
    return ok


def main():
    stages = sys.argv[1:] or ["gemm", "stats", "grad", "critic"]
    cgs = [int(c) for c in os.environ.get("BRINGUP_CG", "1,2").split(",")]
    fns = {"gemm": stage_gemm, "stats": stage_stats, "grad": stage_grad, "critic": stage_critic}
    allok = True
    for cg in cgs:
        for st in stages:
            t0 = time.time()
            try:
                ok = fns[st](cg)
            except Exception as exc:  # keep going: later stages may still tell us something
                ok = False
                print(f"  {st} cg={cg}: EXCEPTION {type(exc).__name__}: {exc}", flush=True)
            print(f"[{st} cg={cg}] {'PASS' if ok else 'FAIL'} ({time.time() - t0:.1f}s)", flush=True)
            allok &= ok
    print("ALL PASS" if allok else "SOME FAILED", flush=True)
    sys.exit(0 if allok else 1)


if __name__ == "__main__":
    main()
