"""Generate tests/golden/head_golden.npz by running the UNMODIFIED reference forward
(`TICA_LapDropout.forward`, python/src/custom_models/models.py:56-82) on CPU in the build
container.  Run from the repo root:  python tests/golden/make_golden.py

What is stored (all small):
  * the three real feature blocks [8,768] (reference test-split rows 0..7 through the
    reference's random-init encoders) and their labels (feature/test_EEG.csv);
  * one replayed noise draw (Laplace [8,2304], Gumbel [2,8,2304]) for `noise_seed`;
  * `w_values`: the reference's only trained-parameter artefact (w_values.txt = sigmoid(DP));
  * per case (eps x hard x DP-kind): reference logits, loss, accuracy, argmax, and the
    reference's autograd gradients for the small tensors (dDP, db1, db2, dWc, dbc) plus
    row slices / sums of the big ones (dW1, dW2);
  * the gate index [8,2304] per DP-kind.
The head weights are NOT stored: they are regenerated from numpy PCG64 (`make_params`).
The script also asserts the oracle restatement is bit-identical to the reference here.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import head_oracle as ho  # noqa: E402
from oracle import ref_shim  # noqa: E402

NOISE_SEED = 1234
PARAM_SEED = 7
D, H, C = 2304, 768, 2


def main():
    torch.set_num_threads(8)
    blocks = ref_shim.real_feature_blocks(8)
    import pandas as pd

    label = torch.tensor(pd.read_csv(os.path.join(ref_shim.REFERENCE_ROOT, "feature/test_EEG.csv"))["label"]
                         .fillna(0).to_numpy()[:8].astype(np.int64)).view(8, 1)
    w_values = np.array(open(os.path.join(ref_shim.REFERENCE_ROOT, "w_values.txt")).read().strip().strip(",").split(","),
                        dtype=np.float32)
    assert w_values.shape == (D,)
    dp_kinds = {"zero": np.zeros(D, np.float32),
                "wvalues": np.log(w_values.astype(np.float64) / (1 - w_values.astype(np.float64))).astype(np.float32)}

    lap, gum = ho.replay_reference_draws(NOISE_SEED, 8, D)
    shim = ref_shim.ShimmedReferenceHead()
    out = dict(eeg=blocks[0].numpy(), act=blocks[1].numpy(), cm=blocks[2].numpy(), label=label.numpy(),
               lap=lap.numpy(), gum=gum.numpy(), w_values=w_values, noise_seed=NOISE_SEED,
               param_seed=PARAM_SEED, torch_version=torch.__version__)
    cases = []
    eps_list = [0.1, 1.0, 8.0, float(np.around(np.float64(0.01), 3))]
    for dp_name, dp in dp_kinds.items():
        p = ho.make_params(D, H, C, seed=PARAM_SEED, dp=dp)
        shim.load(p)
        for eps in eps_list:
            for hard in (True, False):
                for t in shim.model.parameters():
                    t.grad = None
                pred = shim.forward(blocks, eps, hard, NOISE_SEED)          # reference forward
                loss, acc, pred_id, _ = ho.cal_loss(pred, label)            # reference cal_loss restated
                loss.backward()
                g = {n: q.grad.detach().clone() for n, q in shim.model.named_parameters()}
                # oracle restatement must reproduce the reference bit-for-bit here
                po = p.clone(requires_grad=True)
                pred_o, aux = ho.head_forward(blocks, po, eps, lap, gum, hard, return_aux=True)
                loss_o, _, _, _ = ho.cal_loss(pred_o, label)
                loss_o.backward()
                assert torch.equal(pred_o, pred), (dp_name, eps, hard, (pred_o - pred).abs().max())
                assert torch.equal(po.DP.grad, g["DP"]) and torch.equal(po.W1.grad, g["fc_layers.0.weight"])
                key = f"{dp_name}_eps{eps}_{'hard' if hard else 'soft'}"
                cases.append(key)
                out[key + "_logits"] = pred.detach().numpy()
                out[key + "_loss"] = np.float32(loss.item())
                out[key + "_acc"] = np.float32(acc.item())
                out[key + "_pred"] = pred_id.numpy()
                out[key + "_dDP"] = g["DP"].numpy()
                out[key + "_db1"] = g["fc_layers.0.bias"].numpy()
                out[key + "_db2"] = g["fc_layers.2.bias"].numpy()
                out[key + "_dWc"] = g["classifier.weight"].numpy()
                out[key + "_dbc"] = g["classifier.bias"].numpy()
                out[key + "_dW1_rows"] = g["fc_layers.0.weight"][:4].numpy()
                out[key + "_dW2_rows"] = g["fc_layers.2.weight"][:4].numpy()
                out[key + "_dW1_colsum"] = g["fc_layers.0.weight"].sum(0).numpy()
                out[key + "_dW2_colsum"] = g["fc_layers.2.weight"].sum(0).numpy()
                out[key + "_perturbed_rows"] = aux["perturbed"][:2].detach().numpy()
                out[f"{dp_name}_gate_index"] = aux["gate_index"].numpy().astype(np.uint8)
    out["cases"] = np.array(cases)
    out["restatement_bitexact_at_generation"] = True
    path = os.path.join(ROOT, "tests", "golden", "head_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB,", len(cases), "cases")


if __name__ == "__main__":
    main()
