"""Generate tests/golden/testsplit_golden.npz: the reference's EVALUATION LOOP (past_acc.py:218-236) over the
whole 601-row test split -- BASELINE configs 1-2 ("same real features, checked against reference logits /
accuracy") -- by running the UNMODIFIED reference forward in the build container.
Run from the repo root:  python tests/golden/make_golden_testsplit.py          (~5 min of CPU)

  * features: the 601 real test-split inputs (feature/EEG/test_bert.pickle, feature/action/test_clip_v2.pickle)
    pushed through the reference's own encoders (random-init BERT: no weights are shipped or downloadable; the
    head only sees whatever [B,2304] arrives).  Stored as float16 (2.8 MB); BOTH sides consume the fp32 upcast of
    exactly these values, so the rounding is part of the fixture, not of the comparison.
  * labels: feature/test_EEG.csv.
  * per eps in {0.1, 1.0, 8.0}: reference logits [601,2] and argmax, batch size 8 in file order (the last batch has
    ONE row, as in the reference where 601 % 8 == 1), noise drawn by the reference itself from
    torch.manual_seed(BASE_SEED + batch_index) so the draws can be replayed (oracle.replay_reference_draws);
    the epoch metrics exactly as the reference forms them (sum of per-batch mean loss / accuracy divided by the number
    of batches; sklearn f1_score(prediction_all, label_all)) and the record text of past_acc.py:232-237.
The head weights are regenerated from numpy PCG64 (oracle.make_params(seed=PARAM_SEED)); DP = logit(w_values.txt),
the reference's only trained artefact.
"""
import os
import pickle
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import head_oracle as ho  # noqa: E402
from oracle import ref_shim  # noqa: E402

BASE_SEED, PARAM_SEED, BS, D = 4000, 7, 8, 2304
EPS = [0.1, 1.0, 8.0]


def features(n):
    models = ref_shim.import_reference_models()
    torch.manual_seed(980616)
    full = models.TICA_LapDropout("bert-base-uncased").eval()
    with open(os.path.join(ref_shim.REFERENCE_ROOT, "feature/action/test_clip_v2.pickle"), "rb") as f:
        act = pickle.load(f)
    with open(os.path.join(ref_shim.REFERENCE_ROOT, "feature/EEG/test_bert.pickle"), "rb") as f:
        eeg = pickle.load(f)
    outs = [[], [], []]
    for lo in range(0, n, 16):
        hi = min(n, lo + 16)
        ids = torch.tensor(np.stack([eeg[i]["input_ids"] for i in range(lo, hi)]))
        am = torch.tensor(np.stack([eeg[i]["attention_mask"] for i in range(lo, hi)]))
        act_img = torch.tensor(act[lo:hi]).unsqueeze(1)
        act_mask = torch.ones(hi - lo, 1, dtype=torch.long)
        with torch.no_grad():                                    # the three encoder calls of models.py:59-68
            seq, pooled = full.bert(input_ids=ids, attention_mask=am, return_dict=False)
            emb = full.visual_encoder(act_img)
            cm = full.multi_head_decoder(tgt=emb.permute(1, 0, 2), memory=seq.permute(1, 0, 2),
                                         tgt_key_padding_mask=act_mask == 0, memory_key_padding_mask=am == 0)
            cm = cm.permute(1, 0, 2).mean(dim=1)
        for o, t in zip(outs, (pooled, emb.squeeze(1), cm)):
            o.append(t)
        print("encoded", hi, "/", n, flush=True)
    return [torch.cat(o).to(torch.float16) for o in outs]


def main():
    import pandas as pd
    from sklearn.metrics import f1_score

    torch.set_num_threads(8)
    label = torch.tensor(pd.read_csv(os.path.join(ref_shim.REFERENCE_ROOT, "feature/test_EEG.csv"))["label"].fillna(0)
                         .to_numpy().astype(np.int64)).view(-1, 1)
    n = label.shape[0]
    assert n == 601
    f16 = features(n)
    blocks = [b.float() for b in f16]
    w_values = np.array(open(os.path.join(ref_shim.REFERENCE_ROOT, "w_values.txt")).read().strip().strip(",").split(","), dtype=np.float64)
    dp = np.log(w_values / (1 - w_values)).astype(np.float32)
    p = ho.make_params(D, 768, 2, seed=PARAM_SEED, dp=dp)
    shim = ref_shim.ShimmedReferenceHead()
    shim.load(p)
    out = dict(eeg=f16[0].numpy(), act=f16[1].numpy(), cm=f16[2].numpy(), label=label.numpy(), base_seed=BASE_SEED,
               param_seed=PARAM_SEED, batch_size=BS, eps=np.array(EPS), torch_version=torch.__version__)
    for eps in EPS:
        logits, preds = [], []
        epoch_loss_val = epoch_acc_val = 0.0
        sample_size_val = 0
        prediction_all, label_all = [], []
        with torch.no_grad():
            for bi, lo in enumerate(range(0, n, BS)):                           # past_acc.py:218-228
                bl = [b[lo:lo + BS] for b in blocks]
                prediction = shim.forward(bl, eps, True, BASE_SEED + bi)         # the reference's own forward + draws
                loss, accuracy, pred_label_id, label_id = ho.cal_loss(prediction, label[lo:lo + BS])
                sample_size_val += 1
                prediction_all.extend(pred_label_id.numpy())
                label_all.extend(label_id.numpy())
                epoch_loss_val += loss.item()
                epoch_acc_val += accuracy.item()
                # the restatement fed with the replayed draws must agree bit for bit
                lap, gum = ho.replay_reference_draws(BASE_SEED + bi, bl[0].shape[0], D)
                assert torch.equal(ho.head_forward(bl, p, eps, lap, gum, True), prediction), (eps, bi)
                logits.append(prediction)
                preds.append(pred_label_id)
        f1 = f1_score(prediction_all, label_all)
        k = f"eps{eps}"
        out[k + "_logits"] = torch.cat(logits).numpy()
        out[k + "_pred"] = torch.cat(preds).numpy()
        out[k + "_val_loss"] = np.float64(epoch_loss_val / sample_size_val)
        out[k + "_val_acc"] = np.float64(epoch_acc_val / sample_size_val)
        out[k + "_f1"] = np.float64(f1)
        print(k, "val loss %.4f acc %.4f f1 %.4f over %d batches" % (out[k + "_val_loss"], out[k + "_val_acc"], f1, sample_size_val))
    out["restatement_bitexact_at_generation"] = True
    path = os.path.join(ROOT, "tests", "golden", "testsplit_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
