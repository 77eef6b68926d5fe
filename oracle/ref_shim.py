"""Run the UNMODIFIED reference class on CPU with injected feature blocks (TEST
INFRASTRUCTURE; only works where /root/reference exists, i.e. in the build container).

`python/src/custom_models/models.py` imports `opacus` (unused by TICA_LapDropout) and calls
`BertModel.from_pretrained` (no network here).  Both are shimmed *outside* the reference
file, which is imported as-is from where it lies; nothing is copied into this repo.  The
three encoders are then replaced by stub modules that return injected [B,768] blocks, so
`TICA_LapDropout.forward` executes its own lines 69-82 (the hot path) on them.

Used by `tests/golden/make_golden.py` to generate the committed golden vectors and by
`tests/test_oracle.py` (skipped when the reference is absent) to re-check the restatement.
"""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("PGF_REFERENCE_ROOT", "/root/reference")
_MODELS_DIR = os.path.join(REFERENCE_ROOT, "python", "src", "custom_models")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_MODELS_DIR, "models.py"))


def import_reference_models(random_init_bert: bool = True):
    """Import the reference's models.py unchanged, with `opacus` stubbed and
    `BertModel.from_pretrained` replaced by a random-init BertModel(BertConfig())."""
    if "opacus" not in sys.modules:
        sys.modules["opacus"] = types.SimpleNamespace(PrivacyEngine=object)
    from transformers import BertConfig, BertModel

    if random_init_bert and not getattr(BertModel, "_pgf_patched", False):
        BertModel.from_pretrained = classmethod(lambda cls, *a, **k: BertModel(BertConfig()))
        BertModel._pgf_patched = True
    if _MODELS_DIR not in sys.path:
        sys.path.insert(0, _MODELS_DIR)
    import models  # the reference file, read in place

    return models


class _StubBert(nn.Module):
    def __init__(self):
        super().__init__()
        self.block = None

    def forward(self, input_ids=None, attention_mask=None, return_dict=False):
        b = self.block
        return b.new_zeros(b.shape[0], 1, b.shape[1]), b


class _StubVisual(nn.Module):
    def __init__(self):
        super().__init__()
        self.block = None

    def forward(self, x):
        return self.block.unsqueeze(1)  # forward() does .squeeze(1)


class _StubDecoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.block = None

    def forward(self, tgt=None, memory=None, tgt_key_padding_mask=None, memory_key_padding_mask=None):
        return self.block.unsqueeze(0)  # forward() does .permute(1,0,2).mean(dim=1): exact for length 1


class ShimmedReferenceHead:
    """TICA_LapDropout with stub encoders; `forward` runs reference lines models.py:56-82."""

    def __init__(self):
        models = import_reference_models()
        cls = models.TICA_LapDropout
        m = cls.__new__(cls)
        nn.Module.__init__(m)
        # the attributes TICA_LapDropout.__init__ (models.py:35-54) creates, minus the encoders
        m.device = torch.device("cpu")
        m.bert, m.visual_encoder, m.multi_head_decoder = _StubBert(), _StubVisual(), _StubDecoder()
        m.fc_layers = nn.Sequential(nn.Linear(3 * 768, 3 * 768), nn.ReLU(), nn.Linear(3 * 768, 768), nn.Tanh())
        m.classifier = nn.Linear(768, 2)
        m.DP = nn.parameter.Parameter(torch.zeros(1, 768 * 3))
        m.noiser = torch.distributions.laplace.Laplace(torch.tensor([0.0]), torch.tensor([1.0]))
        self.model = m.eval()

    def load(self, p):
        sd = {"fc_layers.0.weight": p.W1, "fc_layers.0.bias": p.b1, "fc_layers.2.weight": p.W2,
              "fc_layers.2.bias": p.b2, "classifier.weight": p.Wc, "classifier.bias": p.bc, "DP": p.DP}
        missing, unexpected = self.model.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected

    def forward(self, blocks, epsilon, hard: bool, seed: int):
        """Seed the global RNG, then call the reference forward: it draws the Laplace
        uniform and the Gumbel exponential itself, in that order."""
        eeg, act, cm = blocks
        m = self.model
        m.bert.block, m.visual_encoder.block, m.multi_head_decoder.block = eeg, act, cm
        B = eeg.shape[0]
        dummy_ids = torch.zeros(B, 4, dtype=torch.long)
        dummy_mask = torch.ones(B, 4, dtype=torch.long)
        torch.manual_seed(seed)
        return m(dummy_ids, dummy_mask, act.new_zeros(B, 1, 512), torch.ones(B, 1, dtype=torch.long), epsilon, hard)


def real_feature_blocks(n_rows: int = 8, seed: int = 980616):
    """[n,768]x3 blocks from the reference's real test-split inputs pushed through the
    reference's own (random-init, eval-mode) encoders.  The head only sees whatever [B,2304]
    arrives, so untrained encoders are fine for a fixture (SURVEY.md section 8c pin (i))."""
    import pickle

    import numpy as np

    models = import_reference_models()
    torch.manual_seed(seed)
    full = models.TICA_LapDropout("bert-base-uncased").eval()
    with open(os.path.join(REFERENCE_ROOT, "feature/action/test_clip_v2.pickle"), "rb") as f:
        act = pickle.load(f)
    with open(os.path.join(REFERENCE_ROOT, "feature/EEG/test_bert.pickle"), "rb") as f:
        eeg = pickle.load(f)
    ids = torch.tensor(np.stack([eeg[i]["input_ids"] for i in range(n_rows)]))
    am = torch.tensor(np.stack([eeg[i]["attention_mask"] for i in range(n_rows)]))
    act_img = torch.tensor(act[:n_rows]).unsqueeze(1)
    act_mask = torch.ones(n_rows, 1, dtype=torch.long)
    with torch.no_grad():
        seq, pooled = full.bert(input_ids=ids, attention_mask=am, return_dict=False)
        emb = full.visual_encoder(act_img)
        cm = full.multi_head_decoder(tgt=emb.permute(1, 0, 2), memory=seq.permute(1, 0, 2),
                                     tgt_key_padding_mask=act_mask == 0, memory_key_padding_mask=am == 0)
        cm = cm.permute(1, 0, 2).mean(dim=1)
    return pooled.contiguous(), emb.squeeze(1).contiguous(), cm.contiguous()


# ----------------------------------------------------------------------------------------------
# the older PriGumbel head (SURVEY.md section 8 row a-alt): train_val.py at the reference root
# ----------------------------------------------------------------------------------------------
def import_reference_train_val():
    """Import the reference's train_val.py unchanged.  Shimmed outside the file: `opacus` (unused by this
    path) and `transformers.AdamW` (removed from transformers 5; imported, never used).  Importing it sets
    CUDA_VISIBLE_DEVICES and seeds the global RNGs (train_val.py:20,35); the env change is undone."""
    if "opacus" not in sys.modules:
        sys.modules["opacus"] = types.SimpleNamespace(PrivacyEngine=object)
    import transformers

    if not hasattr(transformers, "AdamW"):
        transformers.AdamW = object
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    saved = os.environ.get("CUDA_VISIBLE_DEVICES")
    import train_val  # the reference file, read in place

    if saved is None:
        os.environ.pop("CUDA_VISIBLE_DEVICES", None)
    else:
        os.environ["CUDA_VISIBLE_DEVICES"] = saved
    return train_val


class ShimmedPriGumbelHead:
    """train_val.ConcatModel with stub encoders: `forward` runs the reference's own lines 144-157
    (gumbel_dropout, Lap_noise included) and `loss` its own loss_function (train_val.py:80-93)."""

    def __init__(self, tau: float, epsilon: float):
        self.tv = tv = import_reference_train_val()
        cls = tv.ConcatModel
        m = cls.__new__(cls)
        nn.Module.__init__(m)
        # the attributes ConcatModel.__init__ (train_val.py:126-140) creates, minus the encoders
        m.bert, m.visual_encoder, m.multi_head_decoder = _StubBert(), _StubVisual(), _StubDecoder()
        m.w = nn.parameter.Parameter(data=torch.rand(768))
        m.dropout = tv.GumbelSoftmaxDropout(tau)
        m.fc1 = nn.Linear(768 * 3, 768 * 3)
        m.fc2 = nn.Linear(768 * 3, 768)
        m.classifier = nn.Linear(768, 2)
        m.epsilon = epsilon
        self.model = m

    def load(self, p):
        sd = {"fc1.weight": p.W1, "fc1.bias": p.b1, "fc2.weight": p.W2, "fc2.bias": p.b2,
              "classifier.weight": p.Wc, "classifier.bias": p.bc, "w": p.w}
        self.model.load_state_dict(sd, strict=True)

    def forward(self, blocks, seed: int, train: bool):
        """Seed the global RNG, then call the reference forward: it draws the Gumbel exponential [768,2] and the
        Laplace uniform [B,1] itself, in that order; train mode = soft gate, eval mode = hard gate."""
        eeg, act, cm = blocks
        m = self.model.train(train)
        m.bert.block, m.visual_encoder.block, m.multi_head_decoder.block = eeg, act, cm
        B = eeg.shape[0]
        torch.manual_seed(seed)
        return m(act.new_zeros(B, 1, 512), torch.ones(B, 1, dtype=torch.long), torch.zeros(B, 4, dtype=torch.long),
                 torch.ones(B, 4, dtype=torch.long))

    def loss(self, prediction, label, alpha: float):
        return self.tv.loss_function(prediction, label, self.model, alpha, self.model.epsilon)
