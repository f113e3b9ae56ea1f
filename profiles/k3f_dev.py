'''Development driver for K3F (fd_cross_attn_fused): phase-by-phase parity against fp32 torch on the
same bf16-rounded operands, the watchdog record, and graph-replayed timings.  Runs each shape in
its own subprocess so a faulting launch cannot take the other cases down.

    python profiles/k3f_dev.py            # all shapes
    python profiles/k3f_dev.py one C N S  # one shape in this process
'''
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = [(320, 4096), (640, 1024), (1280, 256), (1280, 64), (320, 1000)]


def reference(x, wq, kv, k_off, v_off, idx, wo, bo, heads, scale):
    import torch
    S, N, C = x.shape
    d = C // heads
    q = (x.float() @ wq.float().t()).bfloat16().float()
    attn = torch.empty(S, N, C, device=x.device)
    for s in range(S):
        rows = kv[idx[s] * 80: idx[s] * 80 + 77].float()
        k = rows[:, k_off:k_off + C].view(77, heads, d)
        v = rows[:, v_off:v_off + C].view(77, heads, d)
        qs = q[s].view(N, heads, d)
        p = torch.softmax(torch.einsum('nhd,thd->hnt', qs, k) * scale, dim=-1)
        attn[s] = torch.einsum('hnt,thd->nhd', p, v).reshape(N, C)
    out = attn.bfloat16().float() @ wo.float().t() + bo.float()
    return attn, out


def one(C, N, S, timing=True):
    import ctypes
    import torch
    from flexdiffuse_b200 import _native
    dev = torch.device('cuda:0')
    torch.manual_seed(C + N)
    heads = 8
    d = C // heads
    scale = d ** -0.5
    x = torch.randn(S, N, C, device=dev).bfloat16()
    wq = (torch.randn(C, C, device=dev) * C ** -0.5).bfloat16()
    wo = (torch.randn(C, C, device=dev) * C ** -0.5).bfloat16()
    bo = (torch.randn(C, device=dev) * 0.1).bfloat16()
    n_ctx = 3
    stride = 2 * C + 64
    k_off, v_off = 32, 32 + C
    kv = torch.randn(n_ctx * 80, stride, device=dev).bfloat16()
    kv.view(n_ctx, 80, stride)[:, 77:] = 0
    idx = torch.tensor([(i * 2 + 1) % n_ctx for i in range(S)], dtype=torch.int32, device=dev)
    attn_ref, out_ref = reference(x, wq, kv, k_off, v_off, idx, wo, bo, heads, scale)
    lib = _native.lib()
    res = dict(C=C, N=N, S=S)

    def run(phases, attn=None):
        lib.fd_debug_set_k3f_phases(phases)
        o, a = _native.cross_attn_fused(x, wq, kv, k_off, v_off, idx, wo, bo, heads, 77, 80, scale,
                                        attn=attn)
        torch.cuda.synchronize()
        return o, a, _native.k3f_status()

    def err(got, want):
        g = got.float()
        return dict(max_abs=(g - want).abs().max().item(),
                    rel_l2=((g - want).norm() / want.norm()).item(),
                    nan=int(torch.isnan(g).sum().item()))

    # attention only (to_q + softmax + PV)
    _, a, st = run(3, attn=torch.zeros_like(x))
    res['p3'] = dict(status=st, **err(a, attn_ref))
    # to_out only, from the reference's attention output
    o, _, st = run(5, attn=attn_ref.bfloat16().contiguous())
    res['p5'] = dict(status=st, **err(o, out_ref))
    # everything
    o, a, st = run(7)
    res['all'] = dict(status=st, attn=err(a, attn_ref), out=err(o, out_ref))
    # phase trace of CTA (0,0,0): ns after its start
    lib.fd_debug_set_k3f_phases(7)
    buf = (ctypes.c_longlong * 64)()
    lib.fd_debug_k3f_trace(None, 1)
    for _ in range(3):
        o_na, _ = _native.cross_attn_fused(x, wq, kv, k_off, v_off, idx, wo, bo, heads, 77, 80, scale,
                                           want_attn=False)
    torch.cuda.synchronize()
    res['noattn'] = dict(status=_native.k3f_status(), **err(o_na, out_ref))
    lib.fd_debug_k3f_trace(buf, 0)
    t = list(buf)
    rel = lambda i: (t[i] - t[0]) if t[i] else None
    res['trace_ns'] = dict(setup=rel(1), q_done=rel(2), conv=rel(3), kv=rel(4),
                           heads=[[rel(8 + 4 * j + e) for e in range(4)] for j in range(320 // d)],
                           stored=rel(48), rendezvous=rel(49), out_full=rel(50), end=rel(51))
    if timing:
        lib.fd_debug_set_k3f_phases(7)
        oo, aa = torch.empty_like(x), (torch.empty_like(x) if C != 320 else None)
        f = lambda: _native.cross_attn_fused(x, wq, kv, k_off, v_off, idx, wo, bo, heads, 77, 80, scale,
                                             attn=aa, out=oo, want_attn=False)
        f()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            f()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10):
                f()
        g.replay()
        torch.cuda.synchronize()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        for _ in range(5):
            g.replay()
        b_.record()
        b_.synchronize()
        us = a_.elapsed_time(b_) / 50 * 1e3
        flops = S * (4 * N * C * C + 4 * N * 77 * C)
        res['us'] = us
        res['tflops'] = flops / us / 1e6
        # the three launches it replaces: cuBLAS to_q, K3, cuBLAS to_out (+ bias)
        def unfused():
            q = torch.nn.functional.linear(x, wq)
            o3 = _native.cross_attn(q, kv, k_off, v_off, idx, heads, 77, 80, scale)
            return torch.nn.functional.linear(o3, wo, bo)
        unfused()
        with torch.cuda.stream(side):
            unfused()
        torch.cuda.current_stream().wait_stream(side)
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2):
            for _ in range(10):
                unfused()
        g2.replay()
        torch.cuda.synchronize()
        a_.record()
        for _ in range(5):
            g2.replay()
        b_.record()
        b_.synchronize()
        res['us_unfused'] = a_.elapsed_time(b_) / 50 * 1e3
    print(json.dumps(res))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == 'one':
        one(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
        return
    for S in (2, 8, 32):
        for C, N in (SHAPES if S < 32 else SHAPES[:3]):
            try:
                r = subprocess.run([sys.executable, __file__, 'one', str(C), str(N), str(S)],
                                   capture_output=True, text=True, timeout=180)
                print(r.stdout.strip() or f'C={C} N={N} S={S}: rc={r.returncode} {r.stderr[-600:]}')
            except subprocess.TimeoutExpired:
                print(f'C={C} N={N} S={S}: TIMEOUT')
            sys.stdout.flush()


if __name__ == '__main__':
    main()
