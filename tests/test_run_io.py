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
        # (PNG -> uint8 -> float and to_rgb_tensor differ by an ulp in places; one ulp can flip an e5m2 rounding of a P16 store)
        assert np.abs(got - ref).max() <= 1e-3


def test_pair_index_batch_reader_and_writer(tmp_path):
    """pivlfn.io on the CPU: pair discovery (both naming schemes), size-grouped pinned-free batches, brightness / contrast through
    PIL's ImageEnhance, threaded .flo writer."""
    import PIL.Image
    import PIL.ImageEnhance
    from pivlfn import io as pio
    rng = np.random.default_rng(2)
    sizes = [(16, 24)] * 3 + [(32, 24)] * 2 + [(16, 24)]
    imgs = []
    for k, (h, w) in enumerate(sizes):
        a, b = rng.integers(0, 255, (h, w), dtype=np.uint8), rng.integers(0, 255, (h, w), dtype=np.uint8)
        _write_png(str(tmp_path / f"q{k}_img1.png"), a)
        _write_png(str(tmp_path / f"q{k}_img2.png"), b)
        imgs.append((a, b))
    idx = pio.PairIndex(str(tmp_path), True)
    assert [p.stem for p in idx] == [f"q{k}" for k in range(6)]
    seq = pio.PairIndex(str(tmp_path), False, count=5)                 # 5 frames (q0_img1, q0_img2, q1_img1, ...) -> 4 pairs
    assert len(seq) == 4 and seq[0].stem == "q0_img1" and seq[0].second.endswith("q0_img2.png")
    reader = pio.BatchReader(idx, 0, len(idx), batch=4, pin=False)
    got = []
    for b in reader:
        assert b.first.dtype == torch.uint8 and b.first.shape == b.second.shape and b.first.shape[3] == 3
        got.append((list(b.stems), tuple(b.first.shape[1:3])))
        for k, stem in enumerate(b.stems):
            j = int(stem[1:])
            assert np.array_equal(b.first[k, :, :, 0].numpy(), imgs[j][0]) and np.array_equal(b.second[k, :, :, 2].numpy(), imgs[j][1])
        reader.release(b)
    reader.close()
    # consecutive pairs of one frame size per batch, at most 4, order preserved
    assert [g[0] for g in got] == [["q0", "q1", "q2"], ["q3", "q4"], ["q5"]] and [g[1] for g in got] == [(16, 24), (32, 24), (16, 24)]
    x = pio.unpack_u8(torch.from_numpy(np.stack([np.repeat(imgs[0][0][:, :, None], 3, 2)])))
    assert x.shape == (1, 3, 16, 24) and torch.equal(x[0, 0], torch.from_numpy(imgs[0][0]).float() / 255)
    # brightness / contrast: PIL's own enhancers (what torchvision's adjust_brightness / adjust_contrast call on PIL images)
    ref = PIL.Image.open(str(tmp_path / "q0_img1.png")).convert("RGB")
    ref = PIL.ImageEnhance.Contrast(PIL.ImageEnhance.Brightness(ref).enhance(1.3)).enhance(0.7)
    assert np.array_equal(pio.decode_rgb(str(tmp_path / "q0_img1.png"), 1.3, 0.7), np.asarray(ref))
    w = pio.FloWriter(str(tmp_path / "out"))
    flows = torch.from_numpy(rng.standard_normal((2, 5, 7, 2)).astype(np.float32))
    w.submit(flows, ["a", "b"])
    assert sorted(os.path.basename(p) for p in w.close()) == ["a_out.flo", "b_out.flo"]
    assert np.array_equal(read_flow(str(tmp_path / "out" / "b_out.flo")), flows[1].numpy())


def test_inference_eval_dataset():
    """InferenceEval (src/datasets.py:491-564): image pairs + ground-truth flow, centre-cropped to a multiple of 64."""
    import pathlib
    import shutil
    import tempfile
    from src.datasets import InferenceEval
    # like the reference, flow files with 'test' anywhere in their PATH are skipped -- pytest's tmp_path contains the test name
    tmp_path = pathlib.Path(tempfile.mkdtemp(prefix="pivlfn_eval_"))
    try:
        _inference_eval_checks(tmp_path, InferenceEval)
    finally:
        shutil.rmtree(tmp_path, ignore_errors=True)


def _inference_eval_checks(tmp_path, InferenceEval):
    rng = np.random.default_rng(3)
    for k in range(2):
        _write_png(str(tmp_path / f"e{k}_img1.png"), rng.integers(0, 255, (70, 130), dtype=np.uint8))
        _write_png(str(tmp_path / f"e{k}_img2.png"), rng.integers(0, 255, (70, 130), dtype=np.uint8))
        write_flow(rng.standard_normal((70, 130, 2)).astype(np.float32), str(tmp_path / f"e{k}_flow.flo"))
    write_flow(np.zeros((70, 130, 2), np.float32), str(tmp_path / "orphan_flow.flo"))       # no images: skipped
    ds = InferenceEval(root=str(tmp_path))
    assert len(ds) == 2 and ds.frame_size == (130, 70)
    # the reference takes the size from PIL's (width, height) = (130, 70) -> [128, 64] and hands it to Crop as (height, width):
    # a 128-row crop of a 70-row frame needs padding, which Crop refuses without a padding colour
    assert ds.render_size == [128, 64]
    with pytest.raises(RuntimeError):
        ds[0]
    sq = tmp_path / "sq"
    sq.mkdir()
    _write_png(str(sq / "f_img1.png"), rng.integers(0, 255, (140, 140), dtype=np.uint8))
    _write_png(str(sq / "f_img2.png"), rng.integers(0, 255, (140, 140), dtype=np.uint8))
    gt = rng.standard_normal((140, 140, 2)).astype(np.float32)
    write_flow(gt, str(sq / "f_flow.flo"))
    ds = InferenceEval(root=str(sq))
    (a, b), (f,) = ds[0]
    assert a.shape == (3, 128, 128) and f.shape == (2, 128, 128) and f.dtype == torch.float32
    assert np.array_equal(f.numpy(), gt[6:134, 6:134].transpose(2, 0, 1))
    ds = InferenceEval(inference_size=(64, 64), root=str(sq))                                   # 140 % 64 != 0 -> recomputed anyway
    assert ds.render_size == [128, 128]


def test_inference_parser_defaults_like_the_reference():
    """Inference.parser (inference.py:202-213): PIL pair -> ToTensor -> estimate.  The reference's default device is 'cpu', where
    its own forward cannot run (the correlation raises NotImplementedError); the drop-in fails the same way."""
    import PIL.Image
    from inference import Inference
    from pivlfn import synth
    from src.models import piv_liteflownet
    i1, i2, _ = synth.particle_pair(64, 64, 3, "uniform")
    im1, im2 = PIL.Image.fromarray(i1).convert("RGB"), PIL.Image.fromarray(i2).convert("RGB")
    net = piv_liteflownet(synth.synthetic_state_dict("piv", 0), 1)
    inf = Inference(net, netname="models/PIV-LiteFlowNet-en.paramOnly")
    assert inf.netname == "PIV-LiteFlowNet-en" and inf.default == os.path.join("./outputs", "PIV-LiteFlowNet-en")
    with pytest.raises(NotImplementedError):
        Inference.parser(net, im1, im2)
    with pytest.raises(AssertionError):
        Inference.parser(net, im1, im2.resize((32, 32)))


@pytest.mark.gpu
def test_inference_parser_on_gpu_matches_estimate():
    import PIL.Image
    from inference import Inference, estimate
    from pivlfn import synth
    from src.models import piv_liteflownet
    i1, i2, _ = synth.particle_pair(64, 96, 4, "rankine")
    net = piv_liteflownet(synth.synthetic_state_dict("piv", 0), 1).to("cuda")
    out = Inference.parser(net, PIL.Image.fromarray(i1).convert("RGB"), PIL.Image.fromarray(i2).convert("RGB"), device="cuda")
    ref = estimate(net, synth.to_rgb_tensor(i1)[None].cuda(), synth.to_rgb_tensor(i2)[None].cuda())
    assert isinstance(out, np.ndarray) and out.shape == (64, 96, 2) and np.array_equal(out, ref)


@pytest.mark.gpu
def test_run_main_brightness_contrast_sweep(tmp_path):
    """run.main (run.py:97-134): consecutive frames under every (brightness, contrast) factor, reference file names; flows equal
    what estimate returns for the PIL-enhanced frames."""
    import PIL.Image
    import PIL.ImageEnhance
    from inference import estimate
    from pivlfn import synth
    from src.models import piv_liteflownet
    import run as run_mod
    src = tmp_path / "in"
    src.mkdir()
    frames = []
    for k in range(3):
        i1, _, _ = synth.particle_pair(64, 64, 70 + k, "uniform")
        _write_png(str(src / f"cam_{k:03d}.png"), i1)
        frames.append(str(src / f"cam_{k:03d}.png"))
    net = piv_liteflownet(synth.synthetic_state_dict("piv", 0), 1).to("cuda")
    n = run_mod.main(net, str(src), str(tmp_path / "out"), mod_factors=((1.0, 1.0), (1.2, 0.8)), batch=2)
    assert n == 4
    names = sorted(os.listdir(tmp_path / "out"))
    assert names == ["cam_100_100_000_out.flo", "cam_100_100_001_out.flo", "cam_120_080_000_out.flo", "cam_120_080_001_out.flo"]

    def enh(path, b, c):
        im = PIL.Image.open(path).convert("RGB")
        im = PIL.ImageEnhance.Contrast(PIL.ImageEnhance.Brightness(im).enhance(b)).enhance(c)
        return torch.from_numpy(np.asarray(im).transpose(2, 0, 1).copy()).float().div(255)[None].cuda()
    ref = estimate(net, enh(frames[1], 1.2, 0.8), enh(frames[2], 1.2, 0.8))
    assert np.abs(read_flow(str(tmp_path / "out" / "cam_120_080_001_out.flo")) - ref).max() <= 1e-3
    with pytest.raises(NotImplementedError):
        run_mod.main_dl(net, str(src), str(tmp_path / "o2"), device="cpu")
