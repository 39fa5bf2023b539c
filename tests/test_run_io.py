"""run.py's data contract on the CPU: .flo wire format, output naming, the Run dataset (pairs and sequences)."""
import os

import numpy as np
import pytest
import torch

from src.datasets import Run
from src.utils_plot import flowname_modifier, read_flow, write_flow


def test_flo_round_trip_and_wire_format(tmp_path):
    flow = np.random.default_rng(0).standard_normal((5, 7, 2)).astype(np.float32)
    fn = str(tmp_path / "a_out.flo")
    write_flow(flow, fn)
    raw = open(fn, "rb").read()
    assert len(raw) == 12 + 5 * 7 * 2 * 4
    assert raw[:4] == b"PIEH"                                       # float32 202021.25
    assert np.frombuffer(raw[4:12], np.int32).tolist() == [7, 5]      # width, height
    assert np.array_equal(read_flow(fn), flow)
    f3 = np.zeros((4, 4, 3), np.float32)
    write_flow(f3, str(tmp_path / "s.flo"))
    assert read_flow(str(tmp_path / "s.flo"), use_stereo=True).shape == (4, 4, 3)
    with pytest.raises(AssertionError):
        write_flow(flow, str(tmp_path / "a.txt"))
    with pytest.raises(AssertionError):
        read_flow(str(tmp_path / "missing.flo"))


def test_flowname_modifier():
    assert flowname_modifier("/x/DNS_turbulence_img1.tif", "/out") == "/out/DNS_turbulence_out.flo"
    assert flowname_modifier("frame_0007", "/out", pair=False) == "/out/frame_0007_out.flo"


def _write_png(path, arr):
    import PIL.Image
    PIL.Image.fromarray(arr).save(path)


def test_run_dataset_pairs_and_sequences(tmp_path):
    rng = np.random.default_rng(1)
    for k in range(3):
        _write_png(str(tmp_path / f"s{k}_img1.png"), rng.integers(0, 255, (16, 24), dtype=np.uint8))
        _write_png(str(tmp_path / f"s{k}_img2.png"), rng.integers(0, 255, (16, 24), dtype=np.uint8))
    ds = Run(str(tmp_path), is_pair=True)
    assert len(ds) == 3 and ds.name_list == ["s0", "s1", "s2"]
    (a, b), name = ds[1]
    assert a.shape == (3, 16, 24) and a.dtype == torch.float32 and 0.0 <= float(a.min()) and float(a.max()) <= 1.0
    assert torch.equal(a[0], a[1]) and name == "s1"              # grayscale -> RGB with identical channels
    seq = Run(str(tmp_path), is_pair=False, n_images=4)           # 4 frames -> 3 sequential pairs
    assert len(seq) == 3
    with pytest.raises(ValueError):
        Run(str(tmp_path / "nope"))


def test_reference_demo_flo_is_readable():
    """The reference's shipped demo output (images/demo/DNS_turbulence_out.flo) parses with this reader (build container)."""
    p = "/root/reference/images/demo/DNS_turbulence_out.flo"
    if not os.path.isfile(p):
        pytest.skip("reference mount absent")
    f = read_flow(p)
    assert f.shape == (256, 256, 2) and np.isfinite(f).all()


@pytest.mark.gpu
def test_main_dl_writes_the_flows_estimate_returns(tmp_path):
    import sys
    from pivlfn import synth
    from inference import estimate
    from src.models import piv_liteflownet
    import run as run_mod
    src = tmp_path / "in"
    src.mkdir()
    pairs = []
    for k in range(5):
        i1, i2, _ = synth.particle_pair(64, 96, 50 + k, "uniform")
        _write_png(str(src / f"p{k}_img1.png"), i1)
        _write_png(str(src / f"p{k}_img2.png"), i2)
        pairs.append((i1, i2))
    net = piv_liteflownet(synth.synthetic_state_dict("piv", 0), 1).to("cuda")
    n = run_mod.main_dl(net, str(src), str(tmp_path / "out"), is_pair=True, batch=2)
    assert n == 5
    for k, (i1, i2) in enumerate(pairs):
        got = read_flow(str(tmp_path / "out" / f"p{k}_out.flo"))
        a = synth.to_rgb_tensor(i1)[None].cuda()
        b = synth.to_rgb_tensor(i2)[None].cuda()
        ref = estimate(net, a, b)
        assert got.shape == (64, 96, 2)
        assert np.abs(got - ref).max() <= 1e-5
