"""run_infer_tile.py

Usage:
  run_infer_tile.py [--gpu=<id>] [--model=<path>] [--nr_inference_workers=<n>] \
            [--nr_post_proc_workers=<n>] [--batch_size=<n>] [--input_dir=<path>] \
            [--output_dir=<path>] [--patch_input_shape=<n>] [--patch_output_shape=<n>]
  run_infer_tile.py (-h | --help)
  run_infer_tile.py --version

Options:
  -h --help                   Show this string.
  --version                   Show version.
  --gpu=<id>                  GPU list. [default: 0]
  --model=<path>              Path to saved checkpoint.
  --nr_inference_workers=<n>  Number of workers during inference. [default: 0]
  --nr_post_proc_workers=<n>  Number of workers during post-processing. [default: 0]
  --batch_size=<n>            Batch size. [default: 10]
  --input_dir=<path>          Path to input data directory. Assumes the files are not nested within directory.
  --output_dir=<path>         Path to output data directory. Will create automtically if doesn't exist. [default: output/]
  --patch_input_shape=<n>     Shape of input patch to the network- Assume square shape. [default: 448]
  --patch_output_shape=<n>    Shape of network output- Assume square shape. [default: 144]

Same flags and defaults as the reference CLI (run_infer_tile.py:1-23); the work runs on the
B200-native engine (cerberus_b200). Differences a user should know:
  * `--gpu=0,1,..` uses every listed GPU like the reference (nn.DataParallel over the visible
    devices, infer/base.py:46) - as one process per GPU: the script re-launches itself under
    `python -m torch.distributed.run --nproc-per-node N`, rank 0 packs the checkpoint and
    NCCL-broadcasts it once, the sorted file list is sharded over the ranks (it can also be
    started under torchrun directly);
  * --nr_inference_workers / --nr_post_proc_workers size an image-decode and a file-writer
    thread pool; patch extraction, the network and the post-processing run on the GPU;
  * precision: environment variable CERB_PRECISION = f16x2 (default: parity mode, logits within
    1e-3 of the fp32 reference) or f16 (throughput mode); see cerberus_b200/infer/base.py.
"""
import os
import subprocess
import sys

import yaml

from cerberus_b200.cli import parse_usage

if __name__ == "__main__":
    args = parse_usage(__doc__, version="CoBi Gland Inference")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    gpu_ids = [g for g in (args["--gpu"] or "").split(",") if g.strip() != ""]
    if world == 1 and len(gpu_ids) > 1:
        # one process per listed GPU (the reference spreads over all visible GPUs in one process)
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=",".join(gpu_ids))
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               "--nproc-per-node", str(len(gpu_ids)), "--master-addr", "127.0.0.1",
               "--master-port", os.environ.get("CERB_MASTER_PORT", "29533"),
               os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd, env=env))
    if gpu_ids and world == 1:
        os.environ["CUDA_VISIBLE_DEVICES"] = args["--gpu"]

    input_dir = args["--input_dir"]
    output_dir = args["--output_dir"]
    if not os.path.exists(output_dir):
        os.makedirs(output_dir)

    run_root_dir = args["--model"]
    checkpoint_path = "%s/weights.tar" % run_root_dir
    with open("%s/settings.yml" % (run_root_dir)) as fptr:
        run_paramset = yaml.full_load(fptr)

    target_list = ["gland", "lumen", "nuclei", "patch-class"]

    run_args = {
        "nr_inference_workers": int(args["--nr_inference_workers"]),
        "nr_post_proc_workers": int(args["--nr_post_proc_workers"]),
        "batch_size": int(args["--batch_size"]),
        "input_dir": input_dir,
        "output_dir": output_dir,
        "patch_input_shape": int(args["--patch_input_shape"]),
        "patch_output_shape": int(args["--patch_output_shape"]),
        "patch_output_overlap": 0,
        "postproc_list": target_list,
    }

    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from cerberus_b200.infer.tile import InferManager

    infer = InferManager(
        checkpoint_path=checkpoint_path,
        decoder_dict=run_paramset["dataset_kwargs"]["req_target_code"],
        model_args=run_paramset["model_kwargs"],
        device=local_rank if world > 1 else 0,
    )
    infer.process_file_list(run_args)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
