"""GPU: the whole hot-path step — eager vs two-stream vs CUDA-graph replay give identical results,
and the host-fed front end returns the same rows."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cfg():
    from unopose_b200.pipeline import HotPathConfig

    return HotPathConfig(n_template=1500, n_fine=600, n_coarse=96, feat_dim=64, n_proposal1=800, n_proposal2=80,
                         pe=((0.2, 16), (0.4, 32)))


KEYS = ("tem_idx", "fps_idx1", "fps_idx2", "pe_q0", "pe_q1", "pe_r0", "pe_r1", "init_R", "init_t", "pred_R", "pred_t",
        "pred_pose_score")


def test_eager_overlap_graph_identical(cuda):
    from unopose_b200.pipeline import GraphedHotPath, run_hot_path, synthetic_inputs

    cfg = _cfg()
    inp = synthetic_inputs(1, 3, cfg, device=cuda)
    torch.manual_seed(3)
    a = run_hot_path(inp, cfg, overlap=False)
    torch.manual_seed(3)
    b = run_hot_path(inp, cfg, overlap=True)
    torch.cuda.synchronize()
    for k in KEYS:
        assert torch.equal(a[k], b[k]), k
    g = GraphedHotPath(inp, cfg)
    o1 = {k: v.clone() for k, v in g.replay().items() if k in KEYS}
    torch.cuda.synchronize()
    # deterministic stages are identical to the eager run; the coarse stage draws new uniforms per replay
    for k in ("tem_idx", "fps_idx1", "fps_idx2", "pe_r0", "pe_r1", "pred_R", "pred_t", "pred_pose_score"):
        assert torch.equal(a[k], o1[k]), k
    o2 = g.replay()
    torch.cuda.synchronize()
    assert torch.equal(o1["pred_R"], o2["pred_R"])
    assert torch.isfinite(o2["init_R"]).all()
    # refill the static inputs in place -> the graph computes the new batch
    inp2 = synthetic_inputs(2, 3, cfg, device=cuda)
    for k, v in inp2.items():
        if not k.startswith("_"):
            inp[k].copy_(v)
    o3 = g.replay()
    ref = run_hot_path(inp2, cfg, overlap=False)
    torch.cuda.synchronize()
    assert torch.equal(o3["pred_R"], ref["pred_R"]) and torch.equal(o3["tem_idx"], ref["tem_idx"])


def test_host_fed_front_end(cuda):
    from unopose_b200.pipeline import HostFedHotPath, run_hot_path, synthetic_inputs, to_device

    cfg = _cfg()
    B = 2
    hosts = [synthetic_inputs(10 + i, B, cfg, pin=True) for i in range(3)]
    fed = HostFedHotPath(cfg, B, cuda)
    fed.stage(0, hosts[0])
    rows = []
    for i in range(3):
        if i + 1 < 3:
            fed.stage((i + 1) % 2, hosts[i + 1])
        rows.append(fed.run(i % 2).clone())
    for i in range(3):
        o = run_hot_path(to_device(hosts[i], cuda, non_blocking=False), cfg, overlap=False)
        exp = torch.cat([o["pred_R"].reshape(B, 9), o["pred_t"], o["pred_pose_score"].unsqueeze(1)], 1).cpu()
        assert torch.equal(rows[i], exp), i
