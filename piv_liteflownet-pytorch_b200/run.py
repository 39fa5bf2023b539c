#!/usr/bin/env python
"""Drop-in for the reference's batch driver ``run.py`` (flags :24-42, ``main`` :97-134, ``main_dl`` :137-168) on the B200 path.

    python run.py --model piv -i DIR [-o OUT] [-p] [-s START] [-n NUM] [-b B1 B2 ..] [-c C1 C2 ..] [--weights FILE] [--batch 8]

Same flags, same output files (``<name>_out.flo``, Middlebury format; for the brightness / contrast sweep
``<name>_<bbb>_<ccc>_<rest>_out.flo`` with the factors in percent, run.py:119-125) as the reference.  What differs is how
the GPU is fed (``pivlfn.io``): decoder threads fill pinned uint8 staging batches ahead of the device, pairs of equal size
run in batches (``--batch``), the flows come back through pinned buffers and the ``.flo`` files are written by worker
threads while the next batch runs.  Under ``torchrun`` every rank takes a contiguous block of the pair list (no collective
on the data path).  ``--no_cuda`` is accepted and refused: this build has no CPU path (the reference's own CPU path does not
exist either -- its correlation raises NotImplementedError, src/correlation.py:339-340).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402
from inference import estimate  # noqa: E402
from pivlfn import io as pio  # noqa: E402
from pivlfn.feeder import Feeder  # noqa: E402
from pivlfn import shard  # noqa: E402
from src.models import hui_liteflownet, piv_liteflownet  # noqa: E402

parser = argparse.ArgumentParser(description='Inferencing script for LiteFlowNet (B200 path)')
parser.add_argument("--start", "-s", type=int, default=0, help="Input image starting index.")
parser.add_argument("--num_images", "-n", type=int, default=-1, help="Number of image(s) to process from the directory.")
parser.add_argument("--is_pair", "-p", action="store_true", help="To check if the input image format is in pair.")
parser.add_argument("--brightness", "-b", default=None, type=float, nargs="+",
                    help="Add brightness factor to modify all the input images (optional).")
parser.add_argument("--contrast", "-c", default=None, type=float, nargs="+",
                    help="Add contrast factor to modify all the input images (optional).")
parser.add_argument("--model", "-m", type=str, choices=["hui", "piv"], default="piv")
parser.add_argument("--version", "-v", type=int, choices=[1, 2], default=1)
parser.add_argument("--input", "-i", default=["./images/demo"], type=str, nargs="+", help="Input images directory(ies).")
parser.add_argument("--output", "-o", default="./results", type=str, help="Main output directory.")
parser.add_argument("--no_cuda", action="store_true")
parser.add_argument("--weights", "-w", default=None, type=str, help="state_dict file (default: models/pretrain_torch/...)")
parser.add_argument("--batch", type=int, default=8, help="pairs per forward pass")


def get_weights(model: str, weights: str = None):
    """run.py:71-83: the bundled state_dict files; ``torch.load`` of a bare OrderedDict."""
    if weights is None:
        name = "Hui-LiteFlowNet.paramOnly" if model == "hui" else "PIV-LiteFlowNet-en.paramOnly"
        weights = os.path.join("models", "pretrain_torch", name)
    if not os.path.isfile(weights):
        raise ValueError(f"Unknown params input! Weight file '{weights}' is not found.")
    return torch.load(weights, map_location="cpu")


def _run_pairs(net, pairs, lo: int, hi: int, writer: pio.FloWriter, device, batch: int, stems=None,
               brightness: float = 1.0, contrast: float = 1.0) -> int:
    """pairs[lo:hi] through the network in batches; flows go to ``writer`` under ``stems[i]`` (default: the pair's stem)."""
    reader = pio.BatchReader(pairs, lo, hi, batch, brightness=brightness, contrast=contrast)
    # uint8 frames are uploaded on a side stream while the previous batch runs, unpacked to [B,3,H,W] fp32 on the device, the
    # flows come back as [B,H,W,2] (the .flo layout) on a third stream into pinned buffers that the writer threads own
    feeder = Feeder(lambda x1, x2: estimate(net, x1, x2, tensor=True), device, unpack=pio.unpack_u8,
                    to_host=lambda f: f.permute(0, 2, 3, 1).contiguous())
    done = lo
    pending = []                                   # (batch, first index) whose staging buffers are still being uploaded

    def hand_over(items):
        for (names, _b), host, landed in items:
            writer.submit(host, names, landed, feeder.recycle)

    try:
        for b in reader:
            n = len(b.stems)
            names = b.stems if stems is None else [stems[k] for k in range(done, done + n)]
            hand_over(feeder.push(b.first, b.second, (names, b)))
            uploaded = torch.cuda.Event()
            uploaded.record(feeder.up)
            pending.append((b, uploaded))
            while pending and (len(pending) > 2 or pending[0][1].query()):
                pb, ev = pending.pop(0)
                ev.synchronize()                    # the staging buffers may be refilled once their upload has finished
                reader.release(pb)
            done += n
        hand_over(feeder.drain())
        for pb, ev in pending:
            ev.synchronize()
            reader.release(pb)
    finally:
        reader.close()
    return done - lo


def main_dl(net, inputdir: str, savedir: str, is_pair: bool = False, start_id: int = 0, num_images: int = -1,
            device: str = "cuda", batch: int = 8, rank: int = 0, world: int = 1) -> int:
    """run.py:137-168.  Returns the number of pairs this rank wrote."""
    if torch.device(device).type != "cuda":
        raise NotImplementedError("pivlfn: CUDA device required (there is no CPU path)")
    pairs = pio.PairIndex(inputdir, is_pair, num_images, start_id)
    lo, hi = shard.pair_range(len(pairs), rank, world)
    print(f"Processing {hi - lo} of {len(pairs)} pairs of images (rank {rank}/{world})...")
    writer = pio.FloWriter(savedir)
    n = _run_pairs(net, pairs, lo, hi, writer, device, batch)
    writer.close()
    return n


def main(net, inputdir: str, savedir: str, start_id: int = 0, num_images: int = -1, device: str = "cuda",
         mod_factors=((1.0, 1.0),), batch: int = 8, rank: int = 0, world: int = 1) -> int:
    """run.py:97-134: every consecutive frame pair of the directory under every (brightness, contrast) factor; output
    ``<frame name up to its last '_'>_<bbb>_<ccc>_<rest of the name>_out.flo``."""
    if torch.device(device).type != "cuda":
        raise NotImplementedError("pivlfn: CUDA device required (there is no CPU path)")
    pairs = pio.PairIndex(inputdir, False, num_images, start_id)
    lo, hi = shard.pair_range(len(pairs), rank, world)
    writer = pio.FloWriter(savedir)
    total = 0
    for brightness, contrast in mod_factors:
        tag = f"{str(int(brightness * 100)).zfill(3)}_{str(int(contrast * 100)).zfill(3)}"
        stems = []
        for p in pairs:
            head, _, tail = os.path.basename(p.first).rpartition("_")          # 'x_img1.tif' -> 'x', 'img1.tif'
            stems.append(os.path.splitext(f"{head}_{tag}_{tail}")[0])
        total += _run_pairs(net, pairs, lo, hi, writer, device, batch, stems=stems, brightness=brightness, contrast=contrast)
    writer.close()
    return total


def cli():
    args = parser.parse_args()
    if args.no_cuda or not torch.cuda.is_available():
        raise SystemExit("run.py: a CUDA device is required -- this build has no CPU path (--no_cuda cannot be honoured)")
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    weights = get_weights(args.model, args.weights)
    net = (piv_liteflownet if args.model == "piv" else hui_liteflownet)(weights, args.version).to("cuda")
    for d in args.input:
        out = os.path.join(args.output, os.path.basename(os.path.normpath(d)), "flow")
        if args.brightness is None and args.contrast is None:
            main_dl(net, d, out, is_pair=args.is_pair, start_id=args.start, num_images=args.num_images, batch=args.batch,
                    rank=rank, world=world)
        else:
            bs = (1.0,) if args.brightness is None else tuple(args.brightness)
            cs = (1.0,) if args.contrast is None else tuple(args.contrast)
            main(net, d, out, start_id=args.start, num_images=args.num_images, mod_factors=tuple((b, c) for b in bs for c in cs),
                 batch=args.batch, rank=rank, world=world)


if __name__ == "__main__":
    cli()
