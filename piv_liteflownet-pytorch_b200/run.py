#!/usr/bin/env python
"""Drop-in for the reference's batch driver ``run.py`` (flags :24-42, ``main_dl`` :137-168) on the B200 path.

    python run.py --model piv -i DIR [-o OUT] [-p] [-s START] [-n NUM] [--weights FILE] [--batch 8]

Same flags and the same output files (``<name>_out.flo``, Middlebury format) as the reference.  Differences that keep
the GPU busy at several hundred pairs/s: pairs of equal size are run in batches (``--batch``), the flows come back
through pinned host buffers, and the ``.flo`` files are written by a small thread pool while the next batch runs.
Under ``torchrun`` every rank takes a contiguous block of the pair list (no collective on the data path).
The brightness / contrast sweep (``-b/-c``, run.py:86-134) needs torchvision's PIL transforms and is not part of the
accelerated path.
"""
import argparse
import os
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402
from inference import estimate  # noqa: E402
from pivlfn import shard  # noqa: E402
from src.datasets import Run  # noqa: E402
from src.models import hui_liteflownet, piv_liteflownet  # noqa: E402
from src.utils_plot import flowname_modifier, write_flow  # noqa: E402

parser = argparse.ArgumentParser(description='Inferencing script for LiteFlowNet (B200 path)')
parser.add_argument("--start", "-s", type=int, default=0, help="Input image starting index.")
parser.add_argument("--num_images", "-n", type=int, default=-1, help="Number of image(s) to process from the directory.")
parser.add_argument("--is_pair", "-p", action="store_true", help="To check if the input image format is in pair.")
parser.add_argument("--model", "-m", type=str, choices=["hui", "piv"], default="piv")
parser.add_argument("--version", "-v", type=int, choices=[1, 2], default=1)
parser.add_argument("--input", "-i", default=["./images/demo"], type=str, nargs="+", help="Input images directory(ies).")
parser.add_argument("--output", "-o", default="./results", type=str, help="Main output directory.")
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


def main_dl(net, inputdir: str, savedir: str, is_pair: bool = False, start_id: int = 0, num_images: int = -1,
            device: str = "cuda", batch: int = 8, rank: int = 0, world: int = 1) -> int:
    """run.py:137-168 with batching, pinned staging and asynchronous .flo writes.  Returns the number of pairs written."""
    os.makedirs(savedir, exist_ok=True)
    ds = Run(root=inputdir, is_pair=is_pair, n_images=num_images, start_at=start_id)
    lo, hi = shard.pair_range(len(ds), rank, world)
    print(f"Processing {hi - lo} of {len(ds)} pairs of images (rank {rank}/{world})...")
    pool = ThreadPoolExecutor(max_workers=4)
    pending = []
    idx = lo
    while idx < hi:
        items, names = [], []
        shape = None
        while idx < hi and len(items) < batch:
            (im1, im2), name = ds[idx]
            if shape is not None and im1.shape != shape:
                break
            shape = im1.shape
            items.append((im1, im2))
            names.append(name)
            idx += 1
        a = torch.stack([x[0] for x in items]).pin_memory().to(device, non_blocking=True)
        b = torch.stack([x[1] for x in items]).pin_memory().to(device, non_blocking=True)
        flow = estimate(net, a, b, tensor=True)                                   # [B,2,H,W]
        host = torch.empty(flow.shape[0], flow.shape[2], flow.shape[3], 2, dtype=torch.float32).pin_memory()
        host.copy_(flow.permute(0, 2, 3, 1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()

        def flush(host=host, names=names, ev=ev):
            ev.synchronize()
            arr = host.numpy()
            for k, name in enumerate(names):
                write_flow(arr[k], flowname_modifier(name, savedir, pair=False))
        pending.append(pool.submit(flush))
    for f in pending:
        f.result()
    pool.shutdown()
    return hi - lo


def main():
    args = parser.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("run.py: a CUDA device is required (there is no CPU path)")
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    weights = get_weights(args.model, args.weights)
    net = (piv_liteflownet if args.model == "piv" else hui_liteflownet)(weights, args.version).to("cuda")
    for d in args.input:
        out = os.path.join(args.output, os.path.basename(os.path.normpath(d)), "flow")
        main_dl(net, d, out, is_pair=args.is_pair, start_id=args.start, num_images=args.num_images, batch=args.batch,
                rank=rank, world=world)


if __name__ == "__main__":
    main()
