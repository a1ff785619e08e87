"""Accuracy of the tensor-core pyramid backward vs the fp32 SIMT backward for the operand-term / de-bias
variants (env PCFA_BWD_TERMS, PCFA_BWD_DEBIAS are read once per process, so each variant is a subprocess).
The gradient pyramid is a realistic one: 12 lookup-backward scatters around a drifting flow."""
import json, os, subprocess, sys

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, ".")
    from pcfa_b200 import _lib
    from pcfa_b200.corr_block import pyramid_layout
    lib = _lib.load(); P = _lib.ptr; s = _lib.stream()
    B, C, H, W, L, R = 1, 256, 55, 128, 4, 4
    g = torch.Generator().manual_seed(0)
    f1 = torch.randn(B, C, H, W, generator=g).cuda(); f2 = torch.randn(B, C, H, W, generator=g).cuda()
    offs, _, _ = pyramid_layout(B, H, W, L)
    gp = torch.zeros(offs[-1], device="cuda")
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    base = torch.stack([xs, ys]).float()[None]
    flow = 4 * torch.randn(B, 2, H, W, generator=g)
    for it in range(12):
        flow = flow + 0.5 * torch.randn(B, 2, H, W, generator=g)
        coords = (base + flow).cuda().contiguous()
        gout = torch.randn(B, L * 81, H, W, generator=g).cuda()
        assert lib.pcfa_corr_lookup_backward(P(gout), P(coords), P(gp), B, H, W, L, R, s) == 0
    wsb = lib.pcfa_corr_pyramid_workspace_bytes(B, C, H, W, L); wsp = torch.empty(wsb, device="cuda", dtype=torch.uint8)
    res = {}
    outs = {}
    for impl in (1, 3):
        g1 = torch.empty_like(f1); g2 = torch.empty_like(f2)
        assert lib.pcfa_corr_pyramid_backward(P(gp), P(f1), P(f2), P(g1), P(g2), P(wsp), wsb, B, C, H, W, L, impl, s) == 0
        torch.cuda.synchronize()
        outs[impl] = (g1.double(), g2.double())
    for i, n in enumerate(("grad_fmap1", "grad_fmap2")):
        a, b = outs[3][i], outs[1][i]
        res[n] = dict(rel_l2=float((a - b).norm() / b.norm()), scale=float((a * b).sum() / (b * b).sum()) - 1.0,
                      nnz_frac=float((gp != 0).float().mean()))
    print(json.dumps(res))
    sys.exit(0)

rows = []
for terms in (2, 1):
    for debias in (0, 1):
        env = dict(os.environ, PCFA_BWD_TERMS=str(terms), PCFA_BWD_DEBIAS=str(debias))
        out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True, timeout=600)
        line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:]
        rows.append(dict(terms=terms, debias=debias, result=line))
        print(rows[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/bwd_precision.json", "w"), indent=1)
