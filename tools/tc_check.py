"""Bring-up diagnostics for the tcgen05 path (run on the GPU box): every GEMM-shaped entry point, single-pass and
3-pass, against a torch fp64 reference computed on the same device.  Each group runs in its own subprocess with
a timeout so that a trapped kernel (dead context) or a hang cannot take the other groups down.

    python tools/tc_check.py            # all groups
    python tools/tc_check.py gemm_nt    # one group, in-process
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GROUPS = ["gemm_nt", "gemm_nn", "gemm_tn", "fprop", "transpose", "wgrad", "trunc"]


def describe(name, got, ref, tol):
    import torch
    err = (got.double() - ref).abs()
    scale = ref.abs().max().clamp_min(1e-30)
    rel = float(err.max() / scale)
    ok = rel < tol
    msg = "%-58s rel %.3e  %s" % (name, rel, "ok" if ok else "FAIL")
    if not ok:
        bad = err > tol * scale
        flat = bad.reshape(-1, bad.shape[-1])
        rows = flat.any(1).nonzero().flatten()
        cols = flat.any(0).nonzero().flatten()
        msg += "\n    bad %.1f%%; rows[%d]: %s ... ; cols[%d]: %s ..." % (
            100.0 * float(bad.double().mean()), rows.numel(), rows[:12].tolist(), cols.numel(), cols[:12].tolist())
        msg += "\n    got[0,:6]=%s\n    ref[0,:6]=%s" % (got.reshape(-1, got.shape[-1])[0, :6].tolist(),
                                                       ref.reshape(-1, ref.shape[-1])[0, :6].tolist())
    print(msg, flush=True)
    return ok


def run_group(group):
    import torch
    import torch.nn.functional as F
    from imitation_from_observation_b200 import _lib, ops
    from imitation_from_observation_b200._lib import IMPL_TC, IMPL_TC3
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(0)

    def rnd(*shape, scale=1.0):
        return (torch.randn(shape, generator=g) * scale).to(dev)

    ok = True
    impls = [(IMPL_TC3, "x3", 2e-5), (IMPL_TC, "x1", 3e-3)]
    if group.startswith("gemm"):
        tA, tB = {"gemm_nt": (False, True), "gemm_nn": (False, False), "gemm_tn": (True, False)}[group]
        for (M, N, K) in [(128, 64, 32), (128, 64, 128), (128, 256, 64), (256, 128, 96), (100, 96, 40),
                          (512, 1024, 1024), (512, 8192, 1024), (64, 1024, 8192)]:
            if tA and M % 32:
                continue
            A = rnd(*((K, M) if tA else (M, K)))
            B = rnd(*((N, K) if tB else (K, N)), scale=K ** -0.5)
            bias = rnd(N, scale=0.1)
            ref = (A.double().t() if tA else A.double()) @ (B.double().t() if tB else B.double())
            refb = F.leaky_relu(ref + bias.double(), 0.2)
            for impl, tag, tol in impls:
                C = ops.gemm(A, B, transA=tA, transB=tB, impl=impl)
                torch.cuda.synchronize()
                ok &= describe("%s %s M%d N%d K%d plain" % (group, tag, M, N, K), C, ref, tol)
                C = ops.gemm(A, B, transA=tA, transB=tB, bias=bias, act=True, impl=impl)
                torch.cuda.synchronize()
                ok &= describe("%s %s M%d N%d K%d bias+lrelu" % (group, tag, M, N, K), C, refb, tol)
                base = rnd(M, N)
                C = base.clone()
                ops.gemm(A, B, out=C, transA=tA, transB=tB, accumulate=True, impl=impl)     # split-K atomics
                torch.cuda.synchronize()
                ok &= describe("%s %s M%d N%d K%d accumulate" % (group, tag, M, N, K), C, ref + base.double(), tol)
    elif group == "fprop":
        from oracle import tf_semantics as O
        for (N, H, W, C, Co, k, st) in [(2, 8, 8, 32, 32, 3, 1), (2, 8, 8, 64, 32, 3, 1), (2, 8, 8, 32, 64, 3, 2),
                                        (2, 32, 32, 64, 128, 5, 2), (3, 16, 16, 128, 256, 5, 2),
                                        (16, 8, 8, 256, 512, 5, 2), (2, 16, 16, 32, 64, 5, 1)]:
            x, w, b = rnd(N, H, W, C), rnd(k, k, C, Co, scale=(k * k * C) ** -0.5), rnd(Co, scale=0.1)
            ref = O.lrelu(O.conv2d(x.double().cpu(), w.double().cpu(), b.double().cpu(), st)).to(dev)
            for impl, tag, tol in impls:
                y = ops.conv2d(x, w, b, stride=st, act=True, impl=impl)
                torch.cuda.synchronize()
                ok &= describe("fprop %s N%d %dx%d C%d Co%d k%d s%d" % (tag, N, H, W, C, Co, k, st), y, ref, tol)
    elif group == "transpose":
        from oracle import tf_semantics as O
        for (N, P, Q, H, W, C1, C2, NB, Cn, k, st) in [(2, 8, 8, 8, 8, 64, 64, 2, 32, 3, 1),
                                                       (8, 4, 4, 8, 8, 64, 64, 8, 64, 5, 2),
                                                       (16, 4, 4, 8, 8, 512, 512, 8, 256, 5, 2),
                                                       (4, 8, 8, 16, 16, 256, 256, 2, 128, 5, 2),
                                                       (2, 16, 16, 32, 32, 128, 128, 1, 64, 5, 2),
                                                       (2, 8, 8, 16, 16, 64, 0, 0, 32, 5, 2)]:
            x1 = rnd(N, P, Q, C1)
            x2 = rnd(NB, P, Q, C2) if C2 else None
            w = rnd(k, k, Cn, C1 + C2, scale=(k * k * (C1 + C2) / st ** 2) ** -0.5)
            b = rnd(Cn, scale=0.1)
            cat = x1 if x2 is None else torch.cat([x1, x2.repeat((N // NB, 1, 1, 1))], 3)
            ref = O.lrelu(O.conv2d_transpose(cat.double().cpu(), w.double().cpu(), b.double().cpu(), (H, W), st)).to(dev)
            for impl, tag, tol in impls:
                y = ops.conv2d_transpose(x1, w, (H, W), bias=b, stride=st, x2=x2, act=True, impl=impl)
                torch.cuda.synchronize()
                ok &= describe("transpose %s N%d %dx%d->%dx%d C%d+%d Cn%d k%d s%d" % (tag, N, P, Q, H, W, C1, C2, Cn, k, st),
                               y, ref, tol)
    elif group == "wgrad":
        from oracle import tf_semantics as O
        for (N, H, W, Ca, Cb1, Cb2, NB, k, st) in [(4, 8, 8, 32, 32, 0, 0, 3, 1), (4, 8, 8, 32, 64, 0, 0, 3, 2),
                                                   (4, 16, 16, 128, 256, 0, 0, 5, 2), (8, 8, 8, 256, 512, 0, 0, 5, 2),
                                                   (4, 8, 8, 256, 512, 512, 2, 5, 2), (16, 32, 32, 64, 128, 0, 0, 5, 2)]:
            P, Q = -(-H // st), -(-W // st)
            big, s1 = rnd(N, H, W, Ca), rnd(N, P, Q, Cb1)
            s2 = rnd(NB, P, Q, Cb2) if Cb2 else None
            small = s1 if s2 is None else torch.cat([s1, s2.repeat((N // NB, 1, 1, 1))], 3)
            wr = torch.zeros((k, k, Ca, Cb1 + Cb2), dtype=torch.float64, requires_grad=True)
            (O.conv2d(big.double().cpu(), wr, None, st) * small.double().cpu()).sum().backward()
            ref = wr.grad.to(dev)
            for impl, tag, tol in impls:
                dw = torch.zeros((k, k, Ca, Cb1 + Cb2), device=dev)
                ops.conv2d_wgrad(big, s1, k, st, small2=s2, dw=dw, impl=impl)
                torch.cuda.synchronize()
                ok &= describe("wgrad %s N%d %dx%d Ca%d Cb%d+%d k%d s%d" % (tag, N, H, W, Ca, Cb1, Cb2, k, st), dw, ref, tol)
    elif group == "trunc":
        # does kind::tf32 ignore the low 13 mantissa bits of fp32 operands?  compare raw vs pre-truncated operands
        M, N, K = 128, 64, 256
        A, B = rnd(M, K), rnd(N, K, scale=K ** -0.5)
        At = (A.view(torch.int32) & ~0x1FFF).view(torch.float32)
        Bt = (B.view(torch.int32) & ~0x1FFF).view(torch.float32)
        c_raw = ops.gemm(A, B, transB=True, impl=IMPL_TC)
        c_tr = ops.gemm(At, Bt, transB=True, impl=IMPL_TC)
        Ar, Br = ops.round_tf32(A), ops.round_tf32(B)
        c_rn = ops.gemm(Ar, Br, transB=True, impl=IMPL_TC)
        torch.cuda.synchronize()
        print("single-pass: raw vs truncated operands: max |diff| = %.3e (0 => hardware truncates)" %
              float((c_raw - c_tr).abs().max()))
        print("single-pass: raw vs rna-rounded operands: max |diff| = %.3e" % float((c_raw - c_rn).abs().max()))
        ref = At.double() @ Bt.double().t()
        print("single-pass truncated vs exact product of truncated operands: rel %.3e" %
              float((c_tr.double() - ref).abs().max() / ref.abs().max()))
    print("GROUP %s %s" % (group, "OK" if ok else "HAS FAILURES"), flush=True)


def main():
    if len(sys.argv) > 1:
        run_group(sys.argv[1])
        return
    for grp in GROUPS:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), grp], timeout=300, stdout=subprocess.PIPE,
                               stderr=subprocess.STDOUT, text=True)
            out = r.stdout
            print(out[-6000:] if len(out) > 6000 else out)
            if r.returncode != 0:
                print("GROUP %s exited with code %d" % (grp, r.returncode))
        except subprocess.TimeoutExpired as e:
            print("GROUP %s TIMED OUT\n%s" % (grp, (e.stdout or "")[-3000:]))


if __name__ == "__main__":
    main()
