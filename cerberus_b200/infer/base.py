"""Mirror of infer/base.py:9-54 (InferManager): loads the model directory's checkpoint and
builds the `run_step` closure — over the CUDA engine instead of nn.DataParallel."""
import os

import torch

from ..engine import Engine

# Precision of the drop-in CLIs. "f16x2" (default) is the PARITY mode: split hi+lo fp16 operands,
# head logits within 1e-3 max-abs of the reference's fp32 forward, so the thresholded instance
# maps reproduce the reference's. "f16" is the THROUGHPUT mode (plain fp16 operands, what
# bench.py times): ~2x faster, logits within ~0.06 of fp32 (closer than `net.half()` is) - outside
# the 1e-3 contract; label maps agree with the reference except near the 0.5 / 0.55 thresholds
# (rates in bench.py's `parity` block). Select with the environment variable CERB_PRECISION (the
# docopt usage strings stay verbatim) or the `precision=` keyword of InferManager.
DEFAULT_PRECISION = "f16x2"


def default_precision():
    p = os.environ.get("CERB_PRECISION", DEFAULT_PRECISION)
    if p not in ("f16", "f16x2"):
        raise ValueError("CERB_PRECISION must be 'f16' or 'f16x2' (got %r)" % p)
    return p


class InferManager(object):
    def __init__(self, **kwargs):
        self.run_step = None
        self.device = 0
        self.precision = default_precision()
        for variable, value in kwargs.items():
            self.__setattr__(variable, value)
        self.__load_model()
        return

    def __load_model(self):
        """infer/base.py:17-54. `checkpoint_path` is a torch.save({"desc": state_dict}) file whose
        keys may carry the DataParallel `module.` prefix; `model_args` is settings.yml's
        model_kwargs (encoder_backbone_name, decoder_kwargs, considered_tasks)."""
        import torch.distributed as dist
        self.rank, self.world_size = 0, 1
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            # one process per GPU: rank 0 reads, folds and packs the checkpoint, the other ranks
            # receive the packed blob by ONE broadcast - the B200-native stand-in
            # for nn.DataParallel's per-forward replicate (infer/base.py:46)
            from ..dist import broadcast_packed_model
            from ..plan import PackedModel
            self.rank, self.world_size = dist.get_rank(), dist.get_world_size()
            packed = None
            if self.rank == 0:
                saved = torch.load(self.checkpoint_path, map_location="cpu")["desc"]
                packed = PackedModel(saved, self.model_args)
            dev = torch.device("cuda", self.device) if dist.get_backend() == "nccl" else torch.device("cpu")
            packed = broadcast_packed_model(packed, self.model_args, self.rank, self.world_size, dev)
            self.engine = Engine(None, None, device=self.device, precision=self.precision, packed=packed)
        else:
            saved = torch.load(self.checkpoint_path, map_location="cpu")["desc"]
            self.engine = Engine(saved, self.model_args, device=self.device, precision=self.precision)
        self.run_step = lambda input_batch, output_shape: self.engine.run_step(
            input_batch, output_shape)
        return
