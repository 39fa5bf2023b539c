"""Drop-in mirror of the reference's ``src`` package for the forward-pass hot path only:
``src.models`` (LiteFlowNet, LiteFlowNet2, hui_liteflownet, piv_liteflownet, backwarp) and
``src.correlation`` (FunctionCorrelation, ModuleCorrelation).  Put this directory's parent
(``piv_liteflownet-pytorch_b200/``) on ``sys.path`` in place of the reference checkout."""
