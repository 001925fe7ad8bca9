"""Ground-truth box extraction on the device (csrc/preprocess.cu) against the reference goldens
(BoundingBoxesGeneratord.converter, utils.py:438-482), the oracle and scipy's labelling."""
import numpy as np
import pytest
import torch

from oracle import ssd3d_oracle as O
from tests.conftest import load_golden
from tests.golden import golden_inputs as GI

pytestmark = pytest.mark.gpu


def _ops():
    from mslesions3d_b200 import ops
    return ops


@pytest.mark.parametrize("name", list(GI.GTBOX_CASES))
@pytest.mark.parametrize("dtype", [torch.float32, torch.uint8])
def test_gt_boxes_match_reference_golden(name, dtype):
    case, gold = GI.GTBOX_CASES[name], load_golden("gtbox.pt")[name]
    segs = GI.gtbox_inputs(case)
    assert float(segs.astype("float64").sum()) == gold["in_sum"]
    boxes, labels = _ops().gt_boxes_from_segmentation(torch.from_numpy(segs).to(dtype).cuda(), case["n_classes"])
    for v in range(segs.shape[0]):
        if gold["boxes"][v] is None:              # the reference raises on an empty volume; empty lists here
            assert boxes[v].shape == (0, 6) and labels[v].shape == (0,)
            continue
        assert torch.equal(boxes[v].cpu(), gold["boxes"][v]), name      # bit-exact, reference order
        assert torch.equal(labels[v].cpu(), gold["labels"][v])


@pytest.mark.parametrize("name", list(GI.GTBOX_INSTANCE_CASES))
@pytest.mark.parametrize("dtype", [torch.float32, torch.uint8])
def test_gt_boxes_instances_match_reference_golden(name, dtype):
    """"instances" mode (utils.py:439-441,483-513) against the reference's converter: bit-exact boxes in class-then-id
    order, ids outside the ranges ignored, split ids boxed as one, plates dropped, empty volume."""
    case, gold = GI.GTBOX_INSTANCE_CASES[name], load_golden("gtbox_instances.pt")[name]
    segs = GI.gtbox_instance_inputs(case)
    if dtype == torch.uint8 and segs.max() > 255:
        pytest.skip("ids beyond uint8")
    boxes, labels = _ops().gt_boxes_from_instances(torch.from_numpy(segs).to(dtype).cuda(), case["thresholds"])
    for v in range(segs.shape[0]):
        if gold["boxes"][v] is None:
            assert boxes[v].shape == (0, 6) and labels[v].shape == (0,)
            continue
        assert torch.equal(boxes[v].cpu(), gold["boxes"][v]), name
        assert torch.equal(labels[v].cpu(), gold["labels"][v])


def test_gt_boxes_instances_many_ids_against_oracle():
    """Hundreds of ids in three ranges (one of them open), batch of 2, ids listed out of order in the volume."""
    rs = np.random.RandomState(9)
    seg = np.zeros((2, 30, 34, 26), dtype=np.float32)
    for v in range(2):
        for i in rs.permutation(np.concatenate([np.arange(10, 140), np.arange(500, 620), np.arange(5000, 5030)])):
            c = [rs.randint(0, d - 3) for d in seg.shape[1:]]
            e = rs.randint(1, 4, size=3)
            seg[v, c[0]:c[0] + e[0], c[1]:c[1] + e[1], c[2]:c[2] + e[2]] = i
    thr = [(500, 1000), (10, 100), (1000, float("inf"))]
    boxes, labels = _ops().gt_boxes_from_instances(torch.from_numpy(seg).cuda(), thr)
    for v in range(2):
        b, l = O.gt_boxes_from_instances(seg[v], thr)
        assert torch.equal(boxes[v].cpu(), b) and torch.equal(labels[v].cpu(), l)
        assert set(l.tolist()) == {1, 2, 3}
    with pytest.raises(NotImplementedError):
        _ops().gt_boxes_from_instances(torch.from_numpy(seg).cuda(), [(1, 100), (50, 200)])


@pytest.mark.parametrize("size,batch", [((96, 96, 96), 4), ((33, 47, 61), 3), ((160, 192, 160), 1)])
def test_gt_boxes_of_generated_volumes(size, batch):
    """Masks of the synthetic generator (touching cubes merge) at BASELINE sizes: same boxes as the host path that
    feeds training (synthetic.boxes_from_mask = scipy.ndimage.label, the reference's own call)."""
    from mslesions3d_b200 import synthetic
    masks = np.stack([synthetic.generate_volume(i, size, (3, 9), (6, 20), 7)[1] for i in range(batch)])
    boxes, labels = _ops().gt_boxes_from_segmentation(torch.from_numpy(masks.astype(np.uint8)).cuda(), 0)
    for v in range(batch):
        want = synthetic.boxes_from_mask(masks[v])
        assert np.array_equal(boxes[v].cpu().numpy(), want)
        assert labels[v].cpu().tolist() == [1] * want.shape[0]


def test_gt_boxes_noise_volume_against_oracle():
    """Thousands of irregular components, two classes plus an ignored value."""
    rs = np.random.RandomState(5)
    seg = (rs.rand(2, 24, 28, 30) < 0.35) * rs.randint(1, 4, size=(2, 24, 28, 30))
    boxes, labels = _ops().gt_boxes_from_segmentation(torch.from_numpy(seg.astype(np.float32)).cuda(), 2, max_boxes=8192)
    for v in range(2):
        b, l = O.gt_boxes_from_segmentation(seg[v].astype(np.float32), 2)
        assert torch.equal(boxes[v].cpu(), b) and torch.equal(labels[v].cpu(), l)


def test_bounding_boxes_generator_transform_interface():
    """utils.BoundingBoxesGeneratord: dictionary transform, single volume and batch, argument checks."""
    from mslesions3d_b200 import synthetic, utils
    mask = synthetic.generate_volume(3, (48, 48, 48), (2, 5), (6, 14), 0)[1]
    want = torch.from_numpy(synthetic.boxes_from_mask(mask))
    gen = utils.BoundingBoxesGeneratord(keys=["seg"], segmentation_mode="classes", n_classes=1)
    d = gen({"seg": torch.from_numpy(mask.astype(np.float32))[None].cuda(), "img": 1})
    assert torch.equal(d["boxes"].cpu(), want) and d["labels"].cpu().tolist() == [1] * want.shape[0] and d["img"] == 1
    batch = torch.from_numpy(np.stack([mask, np.zeros_like(mask)]).astype(np.float32))[:, None].cuda()
    d = utils.BoundingBoxesGeneratord(keys="seg", segmentation_mode="binary")({"seg": batch})
    assert torch.equal(d["boxes"][0].cpu(), want) and d["boxes"][1].shape == (0, 6)
    with pytest.raises(AssertionError):            # utils.py:417: instances mode needs thresholds
        utils.BoundingBoxesGeneratord(keys=["seg"], segmentation_mode="instances")
    inst = utils.BoundingBoxesGeneratord(keys=["seg"], segmentation_mode="instances", thresholds=[(1, 100)])
    d = inst({"seg": torch.from_numpy(mask.astype(np.float32))[None].cuda() * 7})      # one id: one box around all cubes
    pos = np.argwhere(mask)
    assert d["labels"].cpu().tolist() == [1]
    assert torch.equal(d["boxes"].cpu()[0], torch.tensor(list(pos.min(0)) + list(pos.max(0)), dtype=torch.float32) / 48)
    with pytest.raises(KeyError):
        gen({"image": batch})
    with pytest.raises(RuntimeError):              # more components than the caller allowed for
        noisy = (torch.rand(1, 16, 16, 16) < 0.3).float().cuda()
        _ops().gt_boxes_from_segmentation(noisy, 0, max_boxes=8)
    with pytest.raises(RuntimeError):
        _ops().gt_boxes_from_segmentation(torch.zeros(1, 8, 8, 8), 0)       # CPU tensor


def test_extracted_boxes_feed_the_training_step():
    """Segmentation -> device GT boxes -> training_step: same loss as with the host-extracted boxes."""
    from mslesions3d_b200 import synthetic
    from mslesions3d_b200.ssd3d import LSSD3D
    size = (64, 64, 64)
    x, hb, hl = synthetic.make_batch(4, 1, size, with_boxes=True)
    masks = np.stack([synthetic.generate_volume(i, size)[1] for i in range(4)])
    db, dl = _ops().gt_boxes_from_segmentation(torch.from_numpy(masks.astype(np.uint8)).cuda(), 1)
    for v in range(4):
        assert np.array_equal(db[v].cpu().numpy(), hb[v])
    model = LSSD3D(n_classes=2, input_channels=1, input_size=size, threshold=[0.1, 0.2])
    model.load_state_dict(O.random_state_dict(1, seed=0))
    model = model.cuda().train()
    img = torch.from_numpy(x).cuda()
    l1 = model.training_step({"img": img, "boxes": db, "labels": dl})["loss"]
    l2 = model.training_step({"img": img, "boxes": [torch.from_numpy(b).cuda() for b in hb],
                              "labels": [torch.from_numpy(l).cuda() for l in hl]})["loss"]
    assert float(l1) == float(l2)
