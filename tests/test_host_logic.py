"""CPU: host-side logic around the hot path -- epoch records in the reference's format
(past_acc.py:218-250), the feature cache / loader (data.py:37-45) and the sweep grid."""
import os

import numpy as np
import pytest
import torch

from eeg_multimodal_b200 import feature_cache as fc
from eeg_multimodal_b200 import parallel, records

# model_dict/newfrac_1.0eps/best_record.txt of the reference, byte for byte (LF line ends)
REFERENCE_RECORD = ("Epochs: 48\n        | Train Loss:  0.014\n        | Train Accuracy:  0.998\n"
                    "        | Val Loss:  0.062\n        | Val Accuracy:  0.987\n        | f_1 Score:  0.990\n")


def test_record_format_matches_reference_file():
    assert records.format_record(48, 0.0141, 0.9979, 0.0624, 0.9871, 0.9899) == REFERENCE_RECORD


def test_binary_f1_matches_sklearn():
    from sklearn.metrics import f1_score

    g = torch.Generator().manual_seed(1)
    for n in (1, 7, 601):
        pred = (torch.rand(n, generator=g) < 0.6).long()
        lab = (torch.rand(n, generator=g) < 0.68).long()
        # the reference passes (prediction, label) -- swapped -- which leaves binary F1 unchanged
        assert abs(records.binary_f1(pred, lab) - f1_score(pred.numpy(), lab.numpy(), zero_division=0)) < 1e-12
    assert records.binary_f1(torch.zeros(5), torch.zeros(5)) == 0.0


def test_epoch_meter_is_unweighted_mean_over_batches():
    """past_acc.py:236: sum of per-batch accuracies / number of batches (the 601-row split ends in a
    1-sample batch which weighs as much as a full one)."""
    m = records.EpochMeter()
    for _ in range(75):
        m.update(0.1, 1.0, torch.ones(8), torch.ones(8))
    m.update(2.0, 0.0, torch.zeros(1), torch.ones(1))
    assert abs(m.acc - 75 / 76) < 1e-12 and abs(m.loss - (7.5 + 2.0) / 76) < 1e-12
    assert abs(m.f1() - 2 * 600 / (2 * 600 + 1)) < 1e-12


def test_record_writer_keeps_best_above_half(tmp_path):
    w = records.RecordWriter(str(tmp_path), "newfrac_1.0eps/")
    saved = []

    def meters(f1_hits):
        tr, va = records.EpochMeter(), records.EpochMeter()
        tr.update(0.5, 0.7)
        pred = torch.tensor([1] * f1_hits + [0] * (10 - f1_hits))
        va.update(0.4, f1_hits / 10, pred, torch.ones(10, dtype=torch.long))
        return tr, va

    assert not w.epoch_end(1, *meters(2), state_dict_fn=lambda: saved.append(1) or {"DP": torch.zeros(1, 4)})  # F1 0.333 <= 0.5
    assert not os.path.exists(w.best)
    assert w.epoch_end(2, *meters(8), state_dict_fn=lambda: {"DP": torch.zeros(1, 4)})
    assert not w.epoch_end(3, *meters(6), state_dict_fn=lambda: {"DP": torch.ones(1, 4)})
    assert open(w.whole).read().count("Epochs:") == 3
    assert open(w.best).read().startswith("Epochs: 2\n")
    assert torch.equal(torch.load(w.ckpt)["DP"], torch.zeros(1, 4))  # torch zip despite the .pickle name


def test_feature_cache_roundtrip_and_loader(tmp_path):
    blocks, labels = fc.synthetic_features(601, dims=(768, 768, 768), seed=3)
    path = str(tmp_path / "feat.npz")
    fc.save_features(path, blocks, labels)
    b2, l2 = fc.load_features(path)
    assert all(torch.equal(a, b) for a, b in zip(blocks, b2)) and torch.equal(labels, l2)
    loader = fc.FeatureLoader(b2, l2, batch_size=8, shuffle=True, seed=5, device="cpu")
    assert len(loader) == 76                       # 75 full batches + the 1-sample tail (kept, like the reference)
    seen, sizes = [], []
    for bl, lab in loader:
        sizes.append(lab.shape[0])
        seen.append(bl[0][:, 0].clone())
        assert [b.shape[1] for b in bl] == [768, 768, 768]
    assert sizes == [8] * 75 + [1]
    got = torch.sort(torch.cat(seen)).values
    assert torch.equal(got, torch.sort(blocks[0][:, 0]).values)   # a permutation of the rows
    first_epoch = torch.cat(seen)
    second_epoch = torch.cat([bl[0][:, 0].clone() for bl, _ in loader])
    assert not torch.equal(first_epoch, second_epoch)             # reshuffled every epoch
    with pytest.raises(ValueError):
        fc.save_features(path, [np.zeros((3, 4)), np.zeros((2, 4))], np.zeros(3))


def test_synthetic_features_follow_survey_8d():
    blocks, labels = fc.synthetic_features(20000, dims=(2048, 512))
    assert [tuple(b.shape) for b in blocks] == [(20000, 2048), (20000, 512)]
    assert 0.0 <= float(blocks[0].min()) and float(blocks[0].max()) < 1.0
    assert abs(float(labels.float().mean()) - 0.66) < 0.02


def test_sweep_grid_and_sharding():
    grid = parallel.sweep_grid([0.1, 1, 3, 5, 8, 10], n_seeds=8)
    assert len(grid) == 48 and grid[0]["eps"] == 0.1 and grid[8]["eps"] == 1.0 and grid[1]["seed"] == 980617
    owned = [parallel.shard_models(48, 8, r) for r in range(8)]
    assert all(len(o) == 6 for o in owned)
    assert sorted(sum(owned, [])) == list(range(48))
    assert parallel.shard_models(5, 8, 7) == []                   # more ranks than models: idle rank
    with pytest.raises(ValueError):
        parallel.shard_models(4, 2, 2)
    for gb, world in ((65536, 8), (601, 4), (7, 8)):
        cuts = [parallel.batch_slice(gb, world, r) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == gb
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
        assert max(hi - lo for lo, hi in cuts) - min(hi - lo for lo, hi in cuts) <= 1


def test_dp_init_variants_follow_the_reference_formulas():
    """past_acc.py:94-103: zeros / block constants (.4,.5,.3) / reversed / blocks + (1 - sigmoid(k*z)) - 0.5."""
    from eeg_multimodal_b200 import variants

    dims = (768, 768, 768)
    assert torch.equal(variants.dp_init(None, dims), torch.zeros(2304))
    ni = variants.dp_init("newinit", dims)
    assert ni[0] == pytest.approx(0.4) and ni[768] == pytest.approx(0.5) and ni[-1] == pytest.approx(0.3)
    assert torch.equal(variants.dp_init("tt", dims), torch.flip(ni, [0]))
    rng = np.random.default_rng(0)
    mean_values = rng.uniform(0.2, 0.8, 2304)
    z = (mean_values - np.mean(mean_values)) / np.std(mean_values)
    for name, k in (("newinit_1", 1.0), ("newinit_k1", 1.0), ("newinit_k3", 3.0), ("feawei", 1.0)):
        w_init = 1 - torch.sigmoid(torch.tensor(k * z, dtype=torch.float32))          # the reference's expression
        want = torch.cat((torch.full((1, 768), 0.4), torch.full((1, 768), 0.5), torch.full((1, 768), 0.3)), dim=1) + w_init.unsqueeze(0) - 0.5
        assert torch.allclose(variants.dp_init(name, dims, mean_values), want.view(-1), atol=1e-7)
    with pytest.raises(ValueError):
        variants.dp_init("newinit_k1", dims)              # needs the feature mean
    with pytest.raises(ValueError):
        variants.dp_init("newinit", (2048, 512))          # block constants are defined for the 3-block layout
    with pytest.raises(ValueError):
        variants.dp_init("nonsense", dims)
    grid = parallel.sweep_grid([1.0], n_seeds=2, variants=("newinit", "tt", None))
    assert len(grid) == 6 and [g["variant"] for g in grid[:3]] == ["newinit", "tt", None]


def test_call_plan_logic_with_a_stub_library(monkeypatch):
    """ops.CallPlan (recorded C-ABI call replay of the launch-bound regimes) without a GPU: per-step increments are
    found by difference, input pointers are substituted by position, 32-bit counters wrap, hooks fire in place, and a
    buffer pointer that moved between the recorded steps is refused."""
    from eeg_multimodal_b200 import _lib, ops

    issued = []

    class Stub:
        def __getattr__(self, name):
            def fn(*args):
                issued.append((name, args))
                return 0
            return fn

    monkeypatch.setattr(_lib, "_lib", Stub())
    # pgf_adam_step(p, g, m, v, shadow, n, step, lr, b1, b2, eps, grad_scale, stream); pgf_fill_zero(p, nbytes, stream)
    def adam(p, step):
        return (("adam",), "pgf_adam_step", (p, 200, 300, 400, None, 1000, step, 1e-3, 0.9, 0.999, 1e-8, 1.0, 77))
    hook = (("hook",), None, ("grad",))
    fill = (("fill",), "pgf_fill_zero", (900, 64, 77))
    rec_a, rec_b = [fill, adam(100, 5), hook], [fill, adam(111, 7), hook]           # recorded two steps apart
    plan = ops.CallPlan(rec_a, rec_b, {100: "params"}, {111: "params"}, steps_apart=2)
    fired = []
    plan.replay(3, {"params": 555}, hook=fired.append)
    assert fired == ["grad"] and [n for n, _ in issued] == ["pgf_fill_zero", "pgf_adam_step"]
    args = issued[1][1]
    assert args[0] == 555 and args[6] == 8 and args[1:6] == (200, 300, 400, None, 1000)   # step 5 + 3 * (7-5)/2
    # a 32-bit Philox offset wraps instead of overflowing the ctypes argument
    sig = _lib.SIGNATURES["pgf_gemm_bf16_ddp"][1]
    base = [1, 8, 2, 8, 1, 4, 4, 4, 42, 0xFFFFFFFE, 0, 3, 4, 64, 5, 0, 77]
    assert len(base) == len(sig)
    nxt = list(base)
    nxt[9] = 0xFFFFFFFF
    plan2 = ops.CallPlan([(("g",), "pgf_gemm_bf16_ddp", tuple(base))], [(("g",), "pgf_gemm_bf16_ddp", tuple(nxt))], {})
    plan2.replay(3, {})
    assert issued[-1][1][9] == 1
    # an internal buffer that moved between the recordings cannot be replayed
    with pytest.raises(RuntimeError, match="pointer argument"):
        ops.CallPlan([adam(100, 5)], [adam(100, 6)[:2] + ((100, 201, 300, 400, None, 1000, 6, 1e-3, 0.9, 0.999, 1e-8, 1.0, 77),)], {})
    with pytest.raises(RuntimeError, match="different call sequences"):
        ops.CallPlan([fill], [fill, fill], {})


def test_call_plan_refuses_pointers_into_caller_inputs(monkeypatch):
    """A pointer argument that lies inside a caller input tensor without being its base (blocks[i] of a [M,B,D] input)
    compares equal between two recorded steps that used the same tensors, but cannot be substituted on replay: refused."""
    from eeg_multimodal_b200 import _lib, ops

    class Stub:
        def __getattr__(self, name):
            return lambda *a: 0

    monkeypatch.setattr(_lib, "_lib", Stub())
    fill = lambda p: (("fill",), "pgf_fill_zero", (p, 64, 77))
    # base pointer 1000 (substituted), slice pointer 1000 + 256 inside the same storage [1000, 2024)
    ok = ops.CallPlan([fill(1000)], [fill(1000)], {1000: "x"}, {1000: "x"}, input_extents=[(1000, 2024)])
    assert ok.calls[0][5] == [(0, "x")]
    with pytest.raises(RuntimeError, match="points into a caller input"):
        ops.CallPlan([fill(1000), fill(1256)], [fill(1000), fill(1256)], {1000: "x"}, {1000: "x"}, input_extents=[(1000, 2024)])


def test_driver_refuses_flags_it_would_reinterpret():
    from eeg_multimodal_b200 import train

    p = train.build_parser()
    train._reject_unsupported(p.parse_args([]))
    train._reject_unsupported(p.parse_args(["--n_dp", "0"]))
    for bad in (["--n_para", "2"], ["--n_dp", "3"], ["--n_class", "3"]):
        with pytest.raises(NotImplementedError):
            train._reject_unsupported(p.parse_args(bad))


def test_reference_results_layout():
    """results.pth keys / shapes of train.py:131-144 for one model: everything a torch.cat over epochs."""
    from eeg_multimodal_b200 import train

    N, ne, D = 13, 5, 24
    ep = lambda: dict(train_loss=torch.rand(40), logits=torch.rand(N, ne, 2), pred=torch.zeros(N, ne, dtype=torch.int64),
                      val_loss=torch.rand(N, ne), Accuracy=torch.rand(ne), F1Score=torch.rand(ne), DP_params=torch.zeros(1, D))
    r = train.reference_results([ep(), {"train_loss": torch.rand(40)}, ep()], ne)
    assert r["logits"].shape == (2 * N, ne, 2) and r["pred"].shape == (2 * N, ne) and r["val_loss"].shape == (2 * N, ne)
    assert r["train_loss"].shape == (120,) and r["Accuracy"].shape == (2 * ne,) and r["DP_params"].shape == (2, D)


def test_resident_dataset_follows_the_host_loader_order():
    """ResidentDataset (device-side gather through one permutation per epoch) yields the batches FeatureLoader / a
    DataLoader(shuffle=True) with the same generator would, incl. the partial last batch (data.py:41-42)."""
    blocks, labels = fc.synthetic_features(21, dims=(8, 4), seed=3)
    a = fc.ResidentDataset(blocks, labels, 8, shuffle=True, seed=5, device="cpu")
    b = fc.FeatureLoader(blocks, labels, 8, shuffle=True, seed=5, device="cpu")
    for _ in range(2):                                           # two epochs: the generator state carries over
        sizes = []
        for (gb, gl), (wb, wl) in zip(a, b):                      # in lockstep: the host loader reuses its staging buffers
            assert torch.equal(gl, wl) and all(torch.equal(x, y) for x, y in zip(gb, wb))
            sizes.append(gl.shape[0])
        assert sizes == [8, 8, 5]
    assert a.n_full == 2


REF = "/root/reference"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "feature/test_EEG.csv")), reason="reference checkout not present")
def test_reference_split_converter_reads_the_shipped_files(tmp_path):
    """The reference's own dataset files (data.py:7-35) -> the tensors its Dataset yields -> feature cache, through a
    stand-in encoder (the real BERT / CLIP-projection / cross-attention stack is out of scope)."""
    args = [os.path.join(REF, p) for p in ("feature/test_EEG.csv", "feature/action/test_clip_v2.pickle", "feature/EEG/test_bert.pickle")]
    raw = fc.read_reference_split(*args)
    assert raw["frame_input"].shape == (601, 1, 512) and raw["frame_input"].dtype == torch.float32
    assert raw["title_input"].shape == (601, 512) and raw["text_mask"].shape == (601, 512) and raw["vedio_mask"].shape == (601, 1)
    assert raw["title_input"][0, 0] == 101 and int(raw["label"].sum()) == 411               # [CLS]; 411 positives of 601
    assert torch.equal(raw["text_mask"], (raw["title_input"] != 0).long())

    def encoder(x):                                              # 4-tuple of data.py -> three [B,768] blocks
        frame, vmask, ids, tmask = x
        g = torch.Generator().manual_seed(0)
        proj = torch.randn(512, 768, generator=g) / 512 ** 0.5
        emb = torch.randn(30522, 768, generator=g)
        pooled = (emb[ids] * tmask[..., None]).sum(1) / tmask.sum(1, keepdim=True)
        vis = frame[:, 0] @ proj
        return pooled, vis, pooled * vis

    out = str(tmp_path / "test_split.npz")
    blocks, labels = fc.convert_reference_split(*args, encoder, out, batch_size=128)
    back, lab = fc.load_features(out)
    assert [tuple(b.shape) for b in back] == [(601, 768)] * 3 and torch.equal(lab, raw["label"])
    assert all(torch.equal(a, b) for a, b in zip(back, blocks))


def test_launch_accounting_follows_the_single_launch_paths():
    """bench.py's `gpu_launches` is a claim: pgf_cls_ce (B <= 8) and pgf_perturb_gate_bwd_dp (B <= 32) finish in their
    main kernel, everything else as tabulated."""
    from eeg_multimodal_b200 import _lib

    ce = [0] * len(_lib.SIGNATURES["pgf_cls_ce"][1])
    bw = [0] * len(_lib.SIGNATURES["pgf_perturb_gate_bwd_dp"][1])
    for B, n_ce, n_bw in ((1, 1, 1), (8, 1, 1), (9, 2, 1), (32, 2, 1), (33, 2, 2), (65536, 2, 2)):
        ce[10], bw[4] = B, B
        assert _lib.launches_of("pgf_cls_ce", ce) == n_ce and _lib.launches_of("pgf_perturb_gate_bwd_dp", bw) == n_bw
    assert _lib.launches_of("pgf_gemm_bf16_ddp", ()) == 2 and _lib.launches_of("pgf_adam_step", ()) == 1


def test_launch_accounting_of_the_many_models_regime():
    """From 24 models per launch on, pgf_linear_bwd_dx is the slab kernel + its finalize launch (csrc/linear_wide.cu)."""
    from eeg_multimodal_b200 import _lib

    args = [0] * 19
    args[12] = 8
    for n_models, want in ((6, 1), (23, 1), (24, 2), (128, 2)):
        args[15] = n_models
        assert _lib.launches_of("pgf_linear_bwd_dx", args) == want
    args[12], args[15] = 601, 48                       # larger batches stay on the ring kernel
    assert _lib.launches_of("pgf_linear_bwd_dx", args) == 1
    assert _lib.launches_of("pgf_memcpy_peer_async", ()) == 0     # a copy-engine transfer, not a kernel


def test_bench_work_model_of_the_round2_kernels():
    """bench.py's algorithmic work per launch for the fp32-parity kernels: six plane-pair products per fp32 multiply-add,
    and split3's bytes (4 read + 6 per plane triple + 4 if the fp32 tensor is written too + 2 for the mask plane)."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    name, bound, work = bench.kernel_work(("gemm_x3", 65536, 2560, 2560, 0, 0, 4))
    assert bound == "tensor" and work == 12.0 * 65536 * 2560 * 2560 and "bf16x3" in name
    _, bound, nbytes = bench.kernel_work(("split3", 65536, 2560, 1, 1, 0, 0))
    assert bound == "hbm" and nbytes == 65536 * 2560 * 10
    _, _, nbytes = bench.kernel_work(("split3", 65536, 2560, 0, 1, 1, 1))
    assert nbytes == 65536 * 2560 * 16
    # one reference step skips the gradients the reference discards: 8 D^2 + 10 D H per sample
    assert bench.flops_per_sample_step(2560, 768) == 8 * 2560 * 2560 + 10 * 2560 * 768


def test_single_process_fanout_and_inactive_overlap_hook():
    """World size 1: the shared-batch upload is a plain copy and the bucketed exchange hook only counts its buckets."""
    from eeg_multimodal_b200 import parallel

    fan = parallel.SharedBatchFanout(6, "cpu")
    dst = [torch.zeros(6, 3), torch.zeros(6, dtype=torch.int64)]
    assert fan.register({0: dst}) == "nccl" and (fan.lo, fan.hi) == (0, 6)
    src = [torch.arange(18.0).view(6, 3), torch.arange(6)]
    fan.upload(0, [fan.host_slice(t) for t in src])
    assert torch.equal(dst[0], src[0]) and torch.equal(dst[1], src[1])
    fan.close()
    hook = parallel.OverlappedAllReduce(w1_chunks=5)
    g = torch.ones(10)
    hook.bucket(g[:4]); hook.bucket(g[4:]); hook.finish(); hook(g)
    assert not hook.active and hook.n_buckets == 2 and torch.equal(g, torch.ones(10))


def test_committed_bench_line_keeps_the_driver_contract():
    """profiles/r2_bench_default.json is the line `python bench.py` printed at HEAD on one B200: every key the driver parses is
    there, with the roofline / cpu_baseline / e2e objects of the measurement contract and the other BASELINE configurations in
    config.also."""
    import json

    root = os.path.dirname(os.path.dirname(__file__))
    d = json.loads(open(os.path.join(root, "profiles", "r2_bench_default.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["metric"] == "model-samples/sec" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["config"]["workload"] == "sweep_synth64k" and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3 and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(d["value"] - 6 * 65536 / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6        # whole-job model-samples/s of the step time
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 65536 * (2560 * 4 + 8) and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.02
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    also = d["config"]["also"]
    assert also["sweep48_b8"]["models_per_gpu"] == 6 and also["sweep48_b8"]["launches_per_step"] == 13.0
    assert also["dp64k"]["scaling"] == "strong" and also["dp64k"]["n_gpus"] == 1
    assert 5.0 < also["fp32x3_synth64k"]["fp32_parity_cost"] < 8.0
    ref = json.loads(open(os.path.join(root, "profiles", "r2_bench_reference_cpu.json")).read().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["metric"] == d["metric"] and ref["unit"] == d["unit"] and ref["gpu_launches"] == 0
    assert ref["config"]["workload"] == d["config"]["workload"] and ref["config"]["batch_per_model"] == d["config"]["batch_per_model"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["cpu_baseline"]["value"] == ref["value"]
