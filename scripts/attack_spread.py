import sys, json, torch
sys.path.insert(0, '.')
torch.backends.cudnn.allow_tf32 = False
from pcfa_b200.adapter import build_network
from pcfa_b200.attack import pcfa_attack
from pcfa_b200.networks.weights import synthetic_pair
net = build_network("RAFT", device="cuda", seed=0, gain=0.5)
i1, i2 = synthetic_pair(0, 128, 160)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    r = pcfa_attack(net, "RAFT", i1.cuda(), i2.cuda(), steps=3, use_graph=bool(rep % 2))
    print(rep, [(round(h["aee_adv_tgt"], 3), round(h["aee_adv_pred"], 3), round(h["l2_delta12"], 5)) for h in r.history])
