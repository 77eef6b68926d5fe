"""Generate tests/golden/prigumbel_golden.npz: the reference's OLDER PriGumbel head (SURVEY.md section 8 row a-alt)
run UNMODIFIED in the build container -- train_val.ConcatModel.forward lines 144-157 (its own gumbel_dropout and
Lap_noise, encoders stubbed so the injected [B,768] blocks arrive at line 150) and train_val.loss_function, with
torch autograd for the gradients and torch.optim.Adam for a 3-step trajectory (train_val.py:178,203-215).
Run from the repo root:  python tests/golden/make_golden_prigumbel.py         (~20 s of CPU)

  * features / labels: rows of tests/golden/testsplit_golden.npz (the real 601-row test split through the reference's
    encoders, stored fp16; both sides consume the fp32 upcast).
  * per case (eps, tau, train/eval, B): logits, total loss, accuracy, predictions from the reference drawing its own
    noise after torch.manual_seed(seed) (replayable: oracle.prigumbel_oracle.replay_reference_draws); for the train-mode
    cases the gradients of w, both biases, the classifier, and Freivalds projections dW @ u, v @ dW of the two big
    weight gradients (u, v from numpy PCG64(PROJ_SEED); the full 2304x2304 gradient would be 21 MB per case).
  * trajectory: 3 reference steps (zero_grad, forward, loss_function, backward(retain_graph=True), Adam.step) at the
    reference's own tau=0.01, eps=1, lr=1e-5 (train_val.py:524-529): per-step loss and the final w, classifier, fc2.bias
    and fp64 projections of the final fc1/fc2 weights.
The head weights come from numpy PCG64 (oracle.prigumbel_oracle.make_params(seed=PARAM_SEED)).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import prigumbel_oracle as po  # noqa: E402
from oracle import ref_shim  # noqa: E402

PARAM_SEED, PROJ_SEED, BASE_SEED, D, H = 11, 5, 7000, 2304, 768
ALPHA = 0.37
#        eps  tau   train  B  row0
CASES = [(1.0, 0.01, True, 8, 0), (1.0, 0.01, False, 8, 8), (0.5, 0.1, True, 8, 16), (4.0, 0.1, False, 8, 24),
         (1.0, 1.0, True, 8, 32), (2.0, 0.01, True, 3, 40), (1.0, 0.01, False, 1, 600)]


def projections(rng_seed):
    rng = np.random.default_rng(rng_seed)
    return {k: torch.tensor(rng.standard_normal(n).astype(np.float32)) for k, n in (("uD", D), ("vD", D), ("vH", H))}


def main():
    torch.set_num_threads(8)
    g = np.load(os.path.join(ROOT, "tests", "golden", "testsplit_golden.npz"))
    blocks_all = [torch.tensor(g[k].astype(np.float32)) for k in ("eeg", "act", "cm")]
    label_all = torch.tensor(g["label"]).view(-1, 1)
    p = po.make_params(D, H, seed=PARAM_SEED)
    pr = projections(PROJ_SEED)
    out = dict(param_seed=PARAM_SEED, proj_seed=PROJ_SEED, base_seed=BASE_SEED, alpha=ALPHA,
               cases=np.array([(e, t, int(tr), b, r0) for e, t, tr, b, r0 in CASES], dtype=np.float64),
               torch_version=torch.__version__)
    for ci, (eps, tau, train, B, r0) in enumerate(CASES):
        shim = ref_shim.ShimmedPriGumbelHead(tau, eps)
        shim.load(p)
        bl = [b[r0:r0 + B] for b in blocks_all]
        label = label_all[r0:r0 + B]
        seed = BASE_SEED + ci
        with torch.set_grad_enabled(train):
            prediction = shim.forward(bl, seed, train)                           # the reference's forward + own draws
            loss, acc, pred, _ = shim.loss(prediction, label, ALPHA)              # the reference's loss_function
        k = f"c{ci}_"
        out[k + "logits"] = prediction.detach().numpy()
        out[k + "loss"] = np.float32(loss.item())
        out[k + "acc"] = np.float32(acc.item())
        out[k + "pred"] = pred.numpy()
        # the restatement fed with the replayed draws must agree bit for bit
        gum, lap = po.replay_reference_draws(seed, B, H, eps)
        q = p.clone(requires_grad=train)
        with torch.set_grad_enabled(train):
            mine = po.head_forward(torch.cat(bl, dim=1), q, tau, not train, gum, lap)
            myloss = po.loss_function(mine, label, q.w, ALPHA, eps)[0]
        assert torch.equal(mine, prediction) and torch.equal(myloss, loss), (ci, (mine - prediction).abs().max())
        if train:
            loss.backward()
            myloss.backward()
            m = shim.model
            ref_g = dict(W1=m.fc1.weight.grad, b1=m.fc1.bias.grad, W2=m.fc2.weight.grad, b2=m.fc2.bias.grad,
                         Wc=m.classifier.weight.grad, bc=m.classifier.bias.grad, w=m.w.grad)
            for name, gr in ref_g.items():
                assert torch.equal(getattr(q, name).grad, gr), (ci, name)
            for name in ("b1", "b2", "Wc", "bc", "w"):
                out[k + "d" + name] = ref_g[name].numpy()
            out[k + "dW1_u"] = (ref_g["W1"].double() @ pr["uD"].double()).numpy()
            out[k + "v_dW1"] = (pr["vD"].double() @ ref_g["W1"].double()).numpy()
            out[k + "dW2_u"] = (ref_g["W2"].double() @ pr["uD"].double()).numpy()
            out[k + "v_dW2"] = (pr["vH"].double() @ ref_g["W2"].double()).numpy()
        print("case", ci, (eps, tau, train, B), "loss %.6f acc %.3f" % (loss.item(), acc.item()), flush=True)

    # ---- 3-step trajectory with the reference's own settings (train_val.py:524-529) ----
    tau, eps, lr, B = 0.01, 1.0, 1e-5, 8
    shim = ref_shim.ShimmedPriGumbelHead(tau, eps)
    shim.load(p)
    opt = torch.optim.Adam(shim.model.parameters(), lr=lr)                       # train_val.py:178
    mine = po.PriGumbelTrainer(p, eps, tau, ALPHA, lr)
    losses = []
    for s in range(3):
        bl = [b[48 + s * B:48 + (s + 1) * B] for b in blocks_all]
        label = label_all[48 + s * B:48 + (s + 1) * B]
        seed = BASE_SEED + 100 + s
        opt.zero_grad()                                                           # :206
        prediction = shim.forward(bl, seed, True)                                 # :208
        loss, acc, _, _ = shim.loss(prediction, label, ALPHA)                     # :210
        loss.backward(retain_graph=True)                                          # :214
        opt.step()                                                                # :215
        losses.append(loss.item())
        gum, lap = po.replay_reference_draws(seed, B, H, eps)
        myl, _, _ = mine.step(torch.cat(bl, dim=1), label, gum, lap)
        assert myl == loss.item(), (s, myl, loss.item())
    m = shim.model
    for name, t in (("W1", m.fc1.weight), ("b1", m.fc1.bias), ("W2", m.fc2.weight), ("b2", m.fc2.bias),
                    ("Wc", m.classifier.weight), ("bc", m.classifier.bias), ("w", m.w)):
        assert torch.equal(getattr(mine.p, name).detach(), t.detach()), name
    out["traj_loss"] = np.array(losses, dtype=np.float32)
    out["traj_w"] = m.w.detach().numpy()
    out["traj_Wc"] = m.classifier.weight.detach().numpy()
    out["traj_bc"] = m.classifier.bias.detach().numpy()
    out["traj_b2"] = m.fc2.bias.detach().numpy()
    out["traj_b1"] = m.fc1.bias.detach().numpy()
    out["traj_W1_u"] = (m.fc1.weight.detach().double() @ pr["uD"].double()).numpy()
    out["traj_W2_u"] = (m.fc2.weight.detach().double() @ pr["uD"].double()).numpy()
    out["restatement_bitexact_at_generation"] = True
    path = os.path.join(ROOT, "tests", "golden", "prigumbel_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e3, "kB")


if __name__ == "__main__":
    main()
