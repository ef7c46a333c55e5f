"""CPU: on-disk formats on either side of the hot path -- autoposeestimation_b200/formats.py."""
import json
import os

import numpy as np
import pytest

from autoposeestimation_b200 import formats, synthetic as synth


def _reference_xyz_parser(path, to_meter=True):
    """Literal restatement of pipeline/utils.py:667-684 (the loop body), used as the checker."""
    input_file = open(path)
    out = []
    while 1:
        input_line = input_file.readline()[1:-2]
        if not input_line:
            break
        input_line = input_line[:-1].split(' ')
        xyz = []
        for number in input_line:
            if number != '':
                xyz.append(float(number) / 1000 if to_meter else float(number))
        out.append([xyz[0], xyz[1], xyz[2]])
    input_file.close()
    return np.array(out)


def test_xyz_round_trip_and_reference_quirk(tmp_path):
    rng = np.random.RandomState(0)
    pts = rng.uniform(-60, 60, size=(200, 3))
    p = str(tmp_path / 'obj.xyz')
    formats.write_xyz(p, pts)
    lines = open(p).read().splitlines()
    assert len(lines) == 200 and all(l.startswith('[') and l.endswith(']') for l in lines)
    assert lines[0] == "%s" % pts[0]                                     # create_pointcloud.py:373-376
    got = formats.read_xyz(p, to_meter=True)
    assert np.array_equal(got, _reference_xyz_parser(p, True))           # same parser behaviour, quirk included
    exact = formats.read_xyz(p, to_meter=False, exact=True)
    assert np.allclose(exact, pts, rtol=0, atol=5e-6 * 60)               # numpy str(): 8 significant digits
    assert np.abs(formats.read_xyz(p, to_meter=False) - pts).max() < 1.0  # the quirk costs at most the last digit of z


def test_meta_json_and_pose_label(tmp_path):
    he = synth.hand_eye()
    rng = np.random.RandomState(1)
    r2e = np.identity(4); r2e[:3, :3] = synth.random_rotation(rng, 1.0); r2e[:3, 3] = rng.uniform(-500, 500, 3)
    obj = np.identity(4); obj[:3, :3] = synth.random_rotation(rng, 2.0)
    intr = dict(width=640, height=480, ppx=320.5, ppy=241.25, fx=615.1, fy=614.9, coeffs=[0.0] * 5)
    meta = formats.frame_meta([0.0] * 6, {'x': 1.0, 'y': 2.0, 'z': 3.0, 'a': 0.1, 'b': 0.2, 'c': 0.3}, obj, r2e, intr, 0.001, False,
                              [float(v) for v in he.reshape(-1)], 7)
    p = str(tmp_path / '000007.meta.json')
    formats.write_json(p, meta)
    assert set(json.load(open(p))) == {'joints', 'pose', 'object_pose', 'robot2endEff_tf', 'intr', 'depth_scale', 'symmetric',
                                       'hand_eye_calibration', 'view_point_id'}               # getData.py:177-221
    m = formats.load_frame_meta(p)
    assert np.array_equal(m['robot2Cam'], np.dot(r2e, he)) and np.array_equal(m['object_rotation'], obj[:3, :3])   # create_pointcloud.py:241-247
    assert m['intr']['fx'] == 615.1 and m['depth_scale'] == 0.001 and m['view_point_id'] == 7
    # pose label algebra, create_labels.py:405-420
    R = synth.random_rotation(rng, 0.5); t = rng.uniform(-100, 100, 3)
    lab = formats.pose_label(he, r2e, R, t, 'mug')
    r2o = np.identity(4); r2o[:3, :3] = R; r2o[:3, 3] = t
    c2o = np.linalg.inv(he) @ np.linalg.inv(r2e) @ r2o
    assert np.allclose(lab['position'], c2o[:3, 3], atol=1e-9) and np.allclose(np.array(lab['rotation']).reshape(3, 3), c2o[:3, :3], atol=1e-12)
    assert lab['cls_name'] == 'mug' and len(lab['cam2robot']) == 16 and len(lab['robot2object']) == 16


def test_frame_batch_loader(tmp_path):
    data, labels = tmp_path / 'data', tmp_path / 'labels'
    data.mkdir(); labels.mkdir()
    he = [float(v) for v in synth.hand_eye().reshape(-1)]
    frames = {}
    for idx in (3, 10):
        fr = synth.render_ellipsoid_frame(idx, H=480, W=640)
        formats.save_png(str(data / '{:06d}.depth.png'.format(idx)), fr['depth'])
        formats.save_png(str(labels / '{:06d}.new_pred.label.png'.format(idx)), fr['label'])
        meta = formats.frame_meta([], {}, np.identity(4), np.identity(4), dict(fr['intr'], width=640, height=480, coeffs=[]), 0.001,
                                  False, he, idx)
        formats.write_json(str(data / '{:06d}.meta.json'.format(idx)), meta)
        frames[idx] = fr
    b = formats.FrameBatchLoader(str(data), str(labels), pin=False).load([3, 10])
    assert b['depth'].shape == (2, 480, 640) and b['label'].dtype.is_floating_point is False
    for k, idx in enumerate((3, 10)):
        assert np.array_equal(b['depth'][k].numpy().view(np.uint16), frames[idx]['depth'])          # 16-bit PNG round trip
        assert np.array_equal(b['label'][k].numpy(), frames[idx]['label'])
        assert np.allclose(b['robot2cam'][k].numpy(), synth.hand_eye())
        assert np.allclose(b['cam'][k].numpy(), [frames[idx]['intr'][q] for q in ('ppx', 'ppy', 'fx', 'fy')])
    with pytest.raises(ValueError):
        formats.load_depth_png(str(labels / '000003.new_pred.label.png'))                          # 8-bit file is not a depth image


def test_xyz_parser_matches_reference_run(tmp_path):
    """The .xyz text written as create_pointcloud.py:373-376 writes it and parsed by the REFERENCE's own parser
    (dataset.py:119-141 = pipeline/utils.py:667-684, run by oracle/gen_golden_geometry.py): read_xyz reproduces the parsed
    model bit for bit (including the dropped last character of unpadded z columns), and write_xyz reproduces the text."""
    import os
    from autoposeestimation_b200 import formats
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'geometry_ref.npz'))
    path = tmp_path / 'ellipsoid.xyz'
    path.write_text(str(g['xyz_text']))
    got = formats.read_xyz(str(path), to_meter=True)
    assert got.shape == g['xyz_parsed_m'].shape and np.array_equal(got, g['xyz_parsed_m'])
    path2 = tmp_path / 'again.xyz'
    formats.write_xyz(str(path2), g['xyz_written'])
    assert path2.read_text() == str(g['xyz_text'])
