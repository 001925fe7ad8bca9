"""On-device synthetic volume generator (generate_artificial_dataset.py:63-105) against its
numpy restatement (oracle/philox_oracle.py): cube lists, masks and raw intensities bit-exact; the full device
pipeline (generate -> NormalizeIntensity(nonzero) -> GT boxes) against the host functions on the same raw data;
distribution checks against the reference generator's construction."""
import numpy as np
import pytest
import torch

from oracle import philox_oracle as PO

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size,channels,n,first_idx,seed,num_objects,object_size", [
    ((32, 32, 32), 1, 3, 0, 0, (1, 5), (6, 14)),
    ((24, 40, 33), 2, 2, 7, 12345678901234, (2, 9), (3, 11)),       # odd voxel count, 64-bit seed, 2 channels
    ((64, 64, 64), 1, 2, 1000000, 3, (1, 5), (6, 14)),
])
def test_generator_bit_exact_vs_numpy_restatement(size, channels, n, first_idx, seed, num_objects, object_size):
    from mslesions3d_b200 import ops
    raw, mask, cubes, n_cubes = ops.generate_volumes(n, channels, size, first_idx, seed, num_objects, object_size)
    raw, mask, cubes, n_cubes = raw.cpu().numpy(), mask.cpu().numpy(), cubes.cpu().numpy(), n_cubes.cpu().numpy()
    for i in range(n):
        want_cubes = PO.cubes_of(seed, first_idx + i, size, num_objects, object_size)
        assert int(n_cubes[i]) == len(want_cubes)
        assert [tuple(int(v) for v in c) for c in cubes[i, :len(want_cubes)]] == want_cubes
        assert num_objects[0] + 1 <= len(want_cubes) <= num_objects[1]
        for side, cd, ch, cw in want_cubes:
            assert object_size[0] <= side < object_size[1]
            assert 0 <= cd <= size[0] - side - 1 and 0 <= ch <= size[1] - side - 1 and 0 <= cw <= size[2] - side - 1
        for c in range(channels):
            want, want_mask = PO.volume(seed, first_idx + i, c, size, want_cubes)
            assert np.array_equal(raw[i, c], want), "volume %d channel %d" % (i, c)
            if c == 0:
                assert np.array_equal(mask[i], want_mask)
    if channels > 1:
        assert not np.array_equal(raw[0, 0], raw[0, 1])              # fresh noise per channel, same cubes
    assert not np.array_equal(raw[0], raw[1])


def test_generator_distribution_matches_the_reference_construction():
    from mslesions3d_b200 import ops
    raw, mask, _, _ = ops.generate_volumes(4, 1, (64, 64, 64), 0, 1)
    x, m = raw.cpu().numpy()[:, 0].astype(np.float64), mask.cpu().numpy().astype(bool)
    bg = x[~m]
    assert abs(bg.mean() - 0.5) < 2e-3 and abs(bg.var() - 1.0 / 12.0) < 1e-3 and 0.0 <= bg.min() and bg.max() < 1.0
    fg = x[m]                                        # clip(U + 0.4, 0, 1): E = 0.4*1 + integral_{0.4}^{1} u du... = 0.82
    assert fg.min() >= 0.4 and fg.max() <= 1.0 and abs(fg.mean() - 0.82) < 2e-2
    assert 0.001 < m.mean() < 0.2


def test_device_pipeline_matches_host_functions_on_the_same_raw_data():
    from mslesions3d_b200 import ops, synthetic
    size = (48, 48, 48)
    vols, boxes, labels = synthetic.make_batch_device(4, 2, size, first_idx=3, random_seed=9, with_boxes=True,
                                                      dtype=torch.float32)
    raw, mask, _, _ = ops.generate_volumes(4, 2, size, 3, 9, (1, 5), synthetic.default_object_size(size))
    raw, mask = raw.cpu().numpy(), mask.cpu().numpy()
    for i in range(4):
        for c in range(2):
            want = synthetic.normalize_nonzero(raw[i, c])
            np.testing.assert_allclose(vols[i, c].cpu().numpy(), want, rtol=2e-5, atol=2e-6)
        want_boxes = synthetic.boxes_from_mask(mask[i])
        assert torch.equal(boxes[i].cpu(), torch.from_numpy(want_boxes))
        assert labels[i].cpu().tolist() == [1] * want_boxes.shape[0]
    bf = synthetic.make_batch_device(4, 2, size, first_idx=3, random_seed=9)
    assert bf.dtype == torch.bfloat16 and torch.equal(bf, vols.to(torch.bfloat16))


def test_generated_batches_train_and_predict():
    """The device-generated batch feeds fit_step / predict_step directly (no host staging at all)."""
    from mslesions3d_b200 import synthetic
    from mslesions3d_b200.ssd3d import LSSD3D
    from oracle import ssd3d_oracle as O
    size = (64, 64, 64)
    model = LSSD3D(n_classes=2, input_channels=1, input_size=size, threshold=[0.1, 0.2], lr=1e-3, min_score=0.3)
    model.load_state_dict(O.random_state_dict(1, seed=2))
    model = model.cuda().train()
    losses = []
    for step in range(3):
        vols, boxes, labels = synthetic.make_batch_device(8, 1, size, first_idx=8 * step, with_boxes=True)
        losses.append(model.fit_step({"img": vols, "boxes": boxes, "labels": labels}).cpu())
    assert all(bool(torch.isfinite(l).all()) for l in losses) and model.fit_skipped_steps() == 0
    model.eval()
    with torch.no_grad():
        b, l, s = model.predict_step({"img": synthetic.make_batch_device(2, 1, size, first_idx=100)}, 0)
    assert len(b) == 2 and all(t.shape[1] == 6 for t in b)
