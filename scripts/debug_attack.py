"""Diagnostic: final adversarial AEE of (a) this package's fused attack loop, (b) the same torch L-BFGS
loop over a plain torch-autograd closure built from the oracle's torch ops, both on the GPU, against
(c) the reference's pcfa_attack on CPU (tests/golden/attack_raft.json)."""
import json, sys
import torch
sys.path.insert(0, '.')
from oracle import torch_ref as TR
from pcfa_b200.adapter import build_network, InputPadder
from pcfa_b200.attack import pcfa_attack, avg_epe
from pcfa_b200.networks.weights import synthetic_pair

torch.backends.cudnn.allow_tf32 = False
gold = json.load(open("tests/golden/attack_raft.json"))
eps, bound, mu = 1e-7, 0.005, 2500. / 0.005
GAIN = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5


def torch_loop(ops, device, steps=3, dtype=torch.float32):
    net = build_network("RAFT", device=device, seed=0, ops=ops, gain=GAIN).to(dtype)
    i1, i2 = synthetic_pair(0, 128, 160)
    i1, i2 = (i1 / 255.).to(device, dtype), (i2 / 255.).to(device, dtype)
    padder = InputPadder(i1.shape)
    i1, i2 = padder.pad(i1, i2)
    w1 = torch.atanh(2. * (1. - eps) * i1 - (1 - eps)).requires_grad_(True)
    w2 = torch.atanh(2. * (1. - eps) * i2 - (1 - eps)).requires_grad_(True)
    opt = torch.optim.LBFGS([w1, w2], max_iter=10)
    target = torch.zeros(1, 2, 128, 160, device=device, dtype=dtype)

    def flow_of():
        x1 = TR.scaled_input(w1, var_change=True, eps_box=eps, make_unit_input=True)
        x2 = TR.scaled_input(w2, var_change=True, eps_box=eps, make_unit_input=True)
        return padder.unpad(net(x1, x2, iters=12, test_mode=True)[1])

    def closure():
        opt.zero_grad()
        d1, d2 = TR.extract_deltas(w1, w2, i1, i2, "change_of_variables", eps_box=eps)
        loss = TR.loss_delta_constraint(flow_of(), target, d1, d2, None, delta_bound=bound, mu=mu, f_type="aee")
        loss.backward()
        return loss
    out = []
    for s in range(steps):
        opt.step(closure)
        with torch.no_grad():
            out.append(float(avg_epe(flow_of(), target)))
    return out

net = build_network("RAFT", device="cuda", seed=0, gain=GAIN)
i1, i2 = synthetic_pair(0, 128, 160)
print("weight gain", GAIN)
for graph in (False, True, True):
    r = pcfa_attack(net, "RAFT", i1.cuda(), i2.cuda(), steps=3, use_graph=graph)
    print("fused loop (graph=%s):" % graph, [h["aee_adv_tgt"] for h in r.history])
print("torch ops on GPU fp32   :", torch_loop(TR, "cuda"))
print("torch ops on GPU (rerun):", torch_loop(TR, "cuda"))
key = "dd_cov_g05_s%d" if GAIN == 0.5 else None
print("reference on CPU (gold) :", [gold[key % k]["aee_adv_tgt"] for k in (1, 2, 3)] if key else gold["dd_cov"]["aee_adv_tgt"])
