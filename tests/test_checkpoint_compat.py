"""CPU, build container only (needs /root/reference): checkpoints move between the unmodified reference model and
pose2room_b200.p2rnet.P2RNet in both directions.

The reference saves `{'net': net.state_dict(), 'optimizer': ..., 'epoch': ..., 'min_loss': ...}` with the model
wrapped in DataParallel / DDP, so every key starts with 'module.' (net_utils/utils.py:57-77), loads it back with
`load_state_dict` on the wrapper (utils.py:141-166) and offers `load_weight` for bare dicts (models/network.py:59-67)."""
import os

import pytest
import torch

from tests import model_helpers as H

pytestmark = pytest.mark.needs_reference


@pytest.mark.parametrize("joints", [25, 53])
def test_checkpoint_round_trip_with_the_reference(tmp_path, joints):
    from oracle import ref_import
    from pose2room_b200.config import P2RConfig
    from pose2room_b200.p2rnet import P2RNet
    ref_net, _ = ref_import.build_reference_model(mode="train", joint_num=joints, num_frames=64, seed=3)
    # what CheckpointIO.save writes for a wrapped model
    wrapped = torch.nn.DataParallel(ref_net) if torch.cuda.is_available() else None
    ref_sd = ref_net.state_dict()
    ckpt = {"net": {"module." + k: v for k, v in ref_sd.items()}, "epoch": 7, "min_loss": 1.25}
    if wrapped is not None:
        assert list(wrapped.state_dict()) == list(ckpt["net"])
    path = os.path.join(str(tmp_path), "model_last.pth")
    torch.save(ckpt, path)

    ours = P2RNet(P2RConfig(mode="train", joint_num=joints, num_frames=64))
    ours_sd = ours.state_dict()
    # same keys in the same order, same shapes, same dtypes (219 entries; gmm_heading mu stays float64)
    assert list(ours_sd) == list(ref_sd) and len(ours_sd) == 219
    for k in ref_sd:
        assert ours_sd[k].shape == ref_sd[k].shape and ours_sd[k].dtype == ref_sd[k].dtype, k
    assert ours_sd["detection.gmm_heading.mdn.mu"].dtype == torch.float64
    # reference checkpoint -> our model, the two ways the reference itself loads one
    loaded = torch.load(path, map_location="cpu")
    ours.load_weight(loaded["net"])                                   # models/network.py:59-67
    for k, v in ours.state_dict().items():
        assert torch.equal(v, ref_sd[k]), k
    ours2 = P2RNet(P2RConfig(mode="train", joint_num=joints, num_frames=64))
    holder = torch.nn.Module()
    holder.module = ours2                                             # what DataParallel / DDP expose
    holder.load_state_dict(loaded["net"])                             # net_utils/utils.py:141-166, strict
    for k, v in ours2.state_dict().items():
        assert torch.equal(v, ref_sd[k]), k
    # our checkpoint -> the reference model, strict
    torch.manual_seed(11)
    with torch.no_grad():
        for p in ours2.parameters():
            p.add_(torch.randn_like(p) * 0.01)
    ref_net.load_state_dict(ours2.state_dict(), strict=True)
    for k, v in ref_net.state_dict().items():
        assert torch.equal(v, ours2.state_dict()[k]), k
    # named children / optimiser specs the reference's trainer relies on (train.py:56-57, optimizers.py:22-39)
    assert [n for n, _ in ours.named_children()] == [n for n, _ in ref_net.named_children()]
    for (n1, c1), (n2, c2) in zip(ours.named_children(), ref_net.named_children()):
        assert set(c1.optim_spec) == set(c2.optim_spec), n1     # values come from the YAML (strings there)
