"""Architecture tables for LiteFlowNet / LiteFlowNet2 (PIV and Hui variants).

Everything here is data: channel counts, kernel sizes, level ranges and the
``state_dict`` key layout.  It restates the constructor arguments found in the
reference (``src/models.py:66-317`` for LiteFlowNet, ``:398-658`` for
LiteFlowNet2, factories ``:719-766``) so that the CUDA model in ``model.py`` has
exactly the reference's parameter names and shapes.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Tuple

PLEVELS = 6
LRELU_SLOPE = 0.1
# index = pyramid level (1..6); index 0 unused
FEAT_CH = [0, 32, 64, 96, 128, 192]            # NetC output channels, 0-based list index = level-1  (src/models.py:70-106)
LEVEL_FEAT_CH = [0, 32, 32, 64, 96, 128, 192]   # NetC channels at level l (level 2 is conv2's 32)  (src/models.py:229)
MATCH_FEAT_CH = [0, 64, 64, 64, 96, 128, 192]   # channels entering M/S after NetC_ext at levels 1,2 (src/models.py:124,198)
KSIZE = [0, 7, 7, 5, 5, 3, 3]                   # flow-head / unfold kernel per level (src/models.py:161,205,225)
SUBPIX_IN = [0, 130, 130, 130, 194, 258, 386]   # (src/models.py:198)
REG_IN = [0, 131, 131, 131, 131, 131, 195]      # (src/models.py:237)
DIST_CH = [0, 49, 49, 25, 25, 9, 9]             # (src/models.py:254)

# NetC conv list: (seq name, index in Sequential, cin, cout, k, stride)  (src/models.py:70-106)
NETC = [
    ("conv1", 0, 3, 32, 7, 1),
    ("conv2", 0, 32, 32, 3, 2), ("conv2", 2, 32, 32, 3, 1), ("conv2", 4, 32, 32, 3, 1),
    ("conv3", 0, 32, 64, 3, 2), ("conv3", 2, 64, 64, 3, 1),
    ("conv4", 0, 64, 96, 3, 2), ("conv4", 2, 96, 96, 3, 1),
    ("conv5", 0, 96, 128, 3, 2),
    ("conv6", 0, 128, 192, 3, 2),
]
# after which NETC entry each level's feature map is complete
NETC_LEVEL_END = {1: 0, 2: 3, 3: 5, 4: 7, 5: 8, 6: 9}

# hidden widths of conv_M / conv_S (without the final 2-channel flow head)
HEAD_V1 = [128, 64, 32]              # src/models.py:154-163
HEAD_V2 = [128, 128, 96, 64, 32]     # src/models.py:487-500
CONV_R = [128, 128, 64, 64, 32, 32]  # src/models.py:236-250


@dataclass(frozen=True)
class ModelCfg:
    name: str
    version: int
    starting_scale: float
    lowest_level: int
    mean: Tuple[float, ...]

    @property
    def levels(self) -> List[int]:
        return list(range(self.lowest_level, PLEVELS + 1))

    @property
    def scalefactor(self) -> List[float]:
        return [self.starting_scale / (2.0 ** l) for l in range(PLEVELS + 1)]

    @property
    def head(self) -> List[int]:
        return HEAD_V1 if self.version == 1 else HEAD_V2

    @property
    def n_ext(self) -> int:
        return len(range(self.lowest_level - 1, 2))  # src/models.py:309-311


HUI_MEAN = (0.411618, 0.434631, 0.454253, 0.410782, 0.433645, 0.452793)
CFGS: Dict[str, ModelCfg] = {
    "piv": ModelCfg("piv", 1, 10.0, 1, (0.173935, 0.180594, 0.192608, 0.172978, 0.179518, 0.191300)),
    "hui": ModelCfg("hui", 1, 40.0, 2, HUI_MEAN),
    "piv2": ModelCfg("piv2", 2, 10.0, 2, (0.194286, 0.190633, 0.191766, 0.194220, 0.190595, 0.191701)),
    "hui2": ModelCfg("hui2", 2, 40.0, 3, HUI_MEAN),
}


def param_specs(cfg: ModelCfg) -> "OrderedDict[str, Tuple[int, ...]]":
    """``state_dict`` keys and shapes in the reference's registration order."""
    sp: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def conv(prefix, cin, cout, kh, kw, bias=True):
        sp[prefix + ".weight"] = (cout, cin, kh, kw)
        if bias:
            sp[prefix + ".bias"] = (cout,)

    for seq, idx, cin, cout, k, _ in NETC:
        conv(f"NetC.{seq}.{idx}", cin, cout, k, k)
    for e in range(cfg.n_ext):
        conv(f"NetC_ext.{e}.conv_ext.0", 32, 64, 1, 1)
    for i, lv in enumerate(cfg.levels):
        p = f"NetE_M.{i}"
        if lv != 6:
            sp[p + ".upConv_M.weight"] = (2, 1, 4, 4)
        if lv < 4:
            sp[p + ".upCorr_M.weight"] = (49, 1, 4, 4)
        cin = 49
        for j, w in enumerate(cfg.head):
            conv(f"{p}.conv_M.{2 * j}", cin, w, 3, 3)
            cin = w
        conv(f"{p}.conv_M.{2 * len(cfg.head)}", cin, 2, KSIZE[lv], KSIZE[lv])
    for i, lv in enumerate(cfg.levels):
        p = f"NetE_S.{i}"
        cin = SUBPIX_IN[lv]
        for j, w in enumerate(cfg.head):
            conv(f"{p}.conv_S.{2 * j}", cin, w, 3, 3)
            cin = w
        conv(f"{p}.conv_S.{2 * len(cfg.head)}", cin, 2, KSIZE[lv], KSIZE[lv])
    for i, lv in enumerate(cfg.levels):
        p = f"NetE_R.{i}"
        K = KSIZE[lv]
        if lv < 5:
            conv(p + ".moduleFeat.0", LEVEL_FEAT_CH[lv], 128, 1, 1)
        cin = REG_IN[lv]
        for j, w in enumerate(CONV_R):
            conv(f"{p}.conv_R.{2 * j}", cin, w, 3, 3)
            cin = w
        if lv < 5:
            conv(p + ".conv_dist_R.0", 32, DIST_CH[lv], K, 1)
            conv(p + ".conv_dist_R.1", DIST_CH[lv], DIST_CH[lv], 1, K)
        else:
            conv(p + ".conv_dist_R.0", 32, DIST_CH[lv], K, K)
        conv(p + ".moduleScaleX", DIST_CH[lv], 1, 1, 1)
        conv(p + ".moduleScaleY", DIST_CH[lv], 1, 1, 1)
    return sp


def conv_flops_per_pixel(cfg: ModelCfg) -> float:
    """2*MAC of every convolution per full-resolution input pixel per pair
    (NetC runs on both images).  SURVEY.md section 8(a): PIV 2 390 317.64, Hui 650 947.64."""
    total = 0.0
    area = {l: 1.0 / (4.0 ** (l - 1)) for l in range(1, 7)}
    lvl = 1
    for seq, idx, cin, cout, k, st in NETC:
        if st == 2:
            lvl += 1
        total += 2 * 2.0 * cin * cout * k * k * area[lvl]
    for name, shp in param_specs(cfg).items():
        if not name.endswith(".weight") or name.startswith("NetC."):
            continue
        cout, cin, kh, kw = shp
        if name.startswith("NetC_ext."):
            e = int(name.split(".")[1])
            # which level uses extension e: list index idx = level-1 uses NetC_ext[(idx-1) % n_ext]
            lv = [l for l in (1, 2) if l >= cfg.lowest_level and ((l - 2) % cfg.n_ext) == e][0]
            total += 2 * 2.0 * cin * cout * area[lv]
            continue
        i = int(name.split(".")[1])
        lv = cfg.levels[i]
        if "upConv_M" in name:
            total += 2.0 * 2 * 4 * area[lv]            # each output px: 2x2 taps, 2 channels
        elif "upCorr_M" in name:
            total += 2.0 * 49 * 4 * area[lv]
        else:
            total += 2.0 * cin * cout * kh * kw * area[lv]
    return total
