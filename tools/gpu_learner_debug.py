"""Developer tool: learner vs torch-autograd restatement, term by term."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_learner as tl
lrn = importlib.import_module("vnl-brax-imitation_b200.learner")
x3 = int(os.environ.get("X3", 1)) == 1
P, V = tl._params(1)
batch, mean, std = tl._batch(2, P, V)
L = lrn.PPOLearner(P, V, tl.T, tl.BM, x3=x3)
L.set_normalizer(mean, std)
m = L.metrics_dict(L.loss_and_grads(batch))
torch.cuda.synchronize()
_, want, gP, gV, aux = lrn.reference_loss(P, V, batch, tl.T, tl.BM, mean, std)
for k in want:
    print("%-20s got %.8f want %.8f diff %.2e" % (k, m[k], want[k], m[k] - want[k]))
print("clip_fraction", m["clip_fraction"])
R = tl.T * tl.BM
rel = tl._rel
print("logits", rel(L.ws["logits"].cpu().numpy(), aux["logits"].cpu().numpy()))
print("baseline", rel(L.ws["val"][:R].cpu().numpy(), aux["baseline"].reshape(-1).cpu().numpy()))
print("vs", rel(L.ws["vs"].cpu().numpy(), aux["vs"].reshape(-1).cpu().numpy()))
print("target_lp", rel(L.ws["target_lp"].cpu().numpy(), aux["target_lp"].reshape(-1).cpu().numpy()))
print("adv (normalised in ref)", rel(((L.ws["adv"] - L.ws["scratch2"][0]) / (L.ws["scratch2"][1] + 1e-8)).cpu().numpy(), aux["adv"].reshape(-1).cpu().numpy()))
print("dlogits", rel(L.ws["dlogits"].cpu().numpy(), aux["dlogits"].cpu().numpy()), "max", float(aux["dlogits"].abs().max()))
d = (L.ws["dlogits"].double() - aux["dlogits"]).abs()
print("  worst rows", torch.topk(d.max(1).values, 5))
print("dval", rel(L.ws["dval"].cpu().numpy(), aux["dbaseline"].reshape(-1).cpu().numpy()))
gp, gv = L.policy_grads(), L.value_grads()
for k, g in gP.items():
    print("policy/%-32s %.2e  (max |g| %.2e)" % (k, rel(gp[k], g.cpu().numpy()), float(g.abs().max())))
for k, g in gV.items():
    print("value/%-33s %.2e  (max |g| %.2e)" % (k, rel(gv[k], g.cpu().numpy()), float(g.abs().max())))
r = int(torch.topk(d.max(1).values, 1).indices[0])
tl_, bl_ = float(L.ws["target_lp"][r]), float(batch["log_prob"][r])
A = float((L.ws["adv"][r] - L.ws["scratch2"][0]) / (L.ws["scratch2"][1] + 1e-8))
print("row", r, "target_lp mine", tl_, "ref", float(aux["target_lp"].reshape(-1)[r]), "behaviour", bl_, "rho mine", np.exp(tl_ - bl_), "ref", float(aux["rho"].reshape(-1)[r]), "A", A, "A ref", float(aux["adv"].reshape(-1)[r]))
print(" mine dlogits", L.ws["dlogits"][r, :6].tolist(), "\n ref", aux["dlogits"][r, :6].tolist())
j = int(torch.argmax(d[r]))
print(" col", j, "mine", float(L.ws["dlogits"][r, j]), "ref", float(aux["dlogits"][r, j]), "logit loc/rs", float(L.ws["logits"][r, j % 30]), float(L.ws["logits"][r, 30 + j % 30]), "raw", float(batch["raw_action"][r, j % 30]))
for name, ref in zip(("h0pre", "h1pre", "d0pre", "d1pre"), aux["pre"]):
    mine = L.ws[name].double()
    flips = ((mine > 0) != (ref > 0))
    print(name, "relu sign flips", int(flips.sum()), "of", ref.numel(), "min |pre| at flips", float(ref[flips].abs().max()) if flips.any() else None,
          "max |pre err|", float((mine - ref).abs().max()))
