"""Mirror of infer/base.py:9-54 (InferManager): loads the model directory's checkpoint and
builds the `run_step` closure — over the CUDA engine instead of nn.DataParallel."""
import torch

from ..engine import Engine


class InferManager(object):
    def __init__(self, **kwargs):
        self.run_step = None
        self.device = 0
        self.precision = "f16"
        for variable, value in kwargs.items():
            self.__setattr__(variable, value)
        self.__load_model()
        return

    def __load_model(self):
        """infer/base.py:17-54. `checkpoint_path` is a torch.save({"desc": state_dict}) file whose
        keys may carry the DataParallel `module.` prefix; `model_args` is settings.yml's
        model_kwargs (encoder_backbone_name, decoder_kwargs, considered_tasks)."""
        saved = torch.load(self.checkpoint_path, map_location="cpu")["desc"]
        self.engine = Engine(saved, self.model_args, device=self.device, precision=self.precision)
        self.run_step = lambda input_batch, output_shape: self.engine.run_step(
            input_batch, output_shape)
        return
