"""The reference's ONE reproducible artefact: ``images/demo/DNS_turbulence_img{1,2}.tif`` -> ``DNS_turbulence_out.flo`` with the
pretrained ``models/pretrain_torch/PIV-LiteFlowNet-en.paramOnly``.  The weight blob is NOT in the mount
(``.MISSING_LARGE_BLOBS``), so these tests skip today; they enable themselves as soon as the file is present under
``/root/reference``, ``baseline/_ref/reference`` (copied by baseline/install_ref.py) or ``$PIVLFN_WEIGHTS_DIR``: then parity is
pinned against an upstream artefact instead of against the reference CODE run here."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.environ.get("PIVLFN_REFERENCE", "/root/reference"), os.path.join(ROOT, "baseline", "_ref", "reference")]


def _find(rel):
    extra = os.environ.get("PIVLFN_WEIGHTS_DIR")
    for base in ([extra] if extra else []) + CANDIDATES:
        for p in (os.path.join(base, rel), os.path.join(base, os.path.basename(rel))):
            if os.path.isfile(p):
                return p
    return None


WEIGHTS = _find("models/pretrain_torch/PIV-LiteFlowNet-en.paramOnly")
IMG1, IMG2 = _find("images/demo/DNS_turbulence_img1.tif"), _find("images/demo/DNS_turbulence_img2.tif")
OUT = _find("images/demo/DNS_turbulence_out.flo")
needs_blob = pytest.mark.skipif(not (WEIGHTS and IMG1 and IMG2 and OUT),
                                reason="pretrained PIV-LiteFlowNet-en.paramOnly absent from the reference mount (.MISSING_LARGE_BLOBS)")


def test_demo_inputs_and_shipped_output_are_readable():
    """What IS shipped parses: the demo pair and the reference's own output for it (256 x 256, finite, multi-pixel flow)."""
    if not (IMG1 and OUT):
        pytest.skip("reference demo files not available (neither the mount nor baseline/_ref/reference)")
    from pivlfn import io as pio
    a = pio.decode_rgb(IMG1)
    f = pio.read_flo(OUT)
    assert a.shape == (256, 256, 3) and f.shape == (256, 256, 2) and np.isfinite(f).all() and np.abs(f).max() > 1.0


@pytest.mark.gpu
@needs_blob
def test_demo_pair_reproduces_the_shipped_flo():
    from inference import estimate
    from pivlfn import io as pio
    from src.models import piv_liteflownet
    net = piv_liteflownet(torch.load(WEIGHTS, map_location="cpu"), 1).to("cuda")
    x1 = pio.unpack_u8(torch.from_numpy(pio.decode_rgb(IMG1))[None].cuda())
    x2 = pio.unpack_u8(torch.from_numpy(pio.decode_rgb(IMG2))[None].cuda())
    flow = estimate(net, x1, x2)
    ref = pio.read_flo(OUT)
    d = np.abs(flow - ref)
    print(f"demo DNS_turbulence vs the shipped .flo: max {d.max():.3e} mean {d.mean():.3e}")
    assert d.max() <= 1e-2 and d.mean() <= 1e-3


@needs_blob
def test_pretrained_blob_loads_strictly():
    from src.models import piv_liteflownet
    net = piv_liteflownet(torch.load(WEIGHTS, map_location="cpu"), 1)
    assert sum(p.numel() for p in net.parameters()) == 6249298          # 252 tensors / 6 249 298 params
