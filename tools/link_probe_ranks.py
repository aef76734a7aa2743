"""What the box's host side gives N ranks copying at the same time (run under torchrun on the GPU box):
1-D pinned copies, and the 2-D shapes the batch pipelines issue (a block-range slice of every stream of the bench batch).
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/link_probe_ranks.py"""
import ctypes as C, json, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aad_b200

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
api, gpu = aad_b200.load()
ctx = gpu.create(local)
gpu.lib.AADGpu_BindHostThread(ctx)
res = {}

def run(name, fn):
    g = (C.c_double * 3)()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    rc = fn(g)
    assert rc == 0, gpu.last_error()
    t = torch.tensor(list(g), dtype=torch.float64, device=dev)
    mine = [round(float(x), 1) for x in t.cpu()]
    if world > 1:
        dist.all_reduce(t)
    res[name] = {"rank0": mine, "sum": [round(float(x), 1) for x in t.cpu()]}

run("1d_2GiB", lambda g: gpu.lib.AADGpu_LinkProbe(ctx, 2 << 30, 2, g))
run("1d_256MiB", lambda g: gpu.lib.AADGpu_LinkProbe(ctx, 256 << 20, 8, g))
run("1d_8MiB", lambda g: gpu.lib.AADGpu_LinkProbe(ctx, 8 << 20, 64, g))
# PCM rows of one of 32 slices of the bench batch: 12,500 clips, 27,562 bytes each, 882,000 bytes apart
run("2d_pcm_slice_12500x27562_pitch882000", lambda g: gpu.lib.AADGpu_LinkProbeRows(ctx, 12500, 27562, 882000, 6, g))
# .aad rows of the same slice: 7,168 bytes each, 224,384 bytes apart
run("2d_aad_slice_12500x7168_pitch224384", lambda g: gpu.lib.AADGpu_LinkProbeRows(ctx, 12500, 7168, 224384, 12, g))
# fewer, longer rows: a quarter of the clips, four slices' worth each
run("2d_3125x110248_pitch882000", lambda g: gpu.lib.AADGpu_LinkProbeRows(ctx, 3125, 110248, 882000, 6, g))
if rank == 0:
    print(json.dumps({"world": world, "order": "[h2d, d2h, both] GB/s", **res}))
gpu.destroy(ctx)
if world > 1:
    dist.destroy_process_group()
