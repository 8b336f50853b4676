"""Multi-GPU check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tools/dist_check.py

Every rank renders its interleaved tiles of the same frame (a) into rank 0's SharedFrame over
NVLink peer stores (the fused resolve+gather, include/g19.h "shared frame") and (b) into a
compact array gathered with NCCL; rank 0 compares both with its own single-rank render of the
whole frame. All three must be bit-identical (RNG keyed on the global pixel).
Covers PATH and REF mode, ragged frame sizes and a frame with fewer tiles than ranks. Exit 0 = ok.
"""
import importlib
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
g19 = importlib.import_module("2019global_b200")
g19dist = importlib.import_module("2019global_b200.dist")
abi = g19.abi

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
stream = torch.cuda.current_stream().cuda_stream
ok = True


def check(name, w, h, mode, scene, **kw):
    global ok
    sc, cam, light = g19.Octree.builtin(scene, w=w, h=h)
    rt = g19.RayTracer(cam, light, device=local)
    rt.setScene(sc)
    rt.start()
    frame = g19dist.shared_frame(rt, w, h)
    for rep in range(3):  # several frames through the same SharedFrame: epochs, release/acquire
        p = rt.params(w, h, mode=mode, rank=rank, world=world, seed=rep, **kw)
        frame.render(p, stream=stream)
        if rank == 0:
            frame.wait(world, stream=stream)
            got_rgb = torch.empty(h * w * 3, dtype=torch.uint8, device=dev)
            got_rad = torch.empty(h * w * 3, dtype=torch.float32, device=dev)
            # copies enqueued behind the wait see the complete frame
            frame.read(rgb=got_rgb.data_ptr(), rad=got_rad.data_ptr(), stream=stream)
            frame.release(stream=stream)
        # (b) NCCL gather of compact tile arrays
        pad = g19dist.padded_len(w, h, world)
        rad_bytes = pad * 12
        payload = torch.zeros(rad_bytes + pad * 3, dtype=torch.uint8, device=dev)
        f_rad = torch.zeros(h * w * 3, dtype=torch.float32, device=dev) if rank == 0 else None
        f_rgb = torch.zeros(h * w * 3, dtype=torch.uint8, device=dev) if rank == 0 else None
        rt.render_tiles(p, t_rad=payload.data_ptr(), t_rgb=payload.data_ptr() + rad_bytes, stream=stream)

        def untile_both(r, buf, _frame):
            rt.untile(w, h, r, world, t_rad=buf.data_ptr(), t_rgb=buf.data_ptr() + rad_bytes, d_rad=f_rad.data_ptr(),
                      d_rgb=f_rgb.data_ptr(), stream=stream)
        g19dist.gather_frame(payload, w, h, 3, f_rgb, untile_both)
        torch.cuda.synchronize()
        if rank == 0:
            one = rt.run(w, h, mode=mode, want=("rgb", "radiance"), seed=rep, **kw)
            ref_rgb = torch.from_numpy(one["rgb"].reshape(-1)).to(dev)
            ref_rad = torch.from_numpy(one["radiance"].reshape(-1)).to(dev)
            a = bool(torch.equal(got_rgb, ref_rgb)) and bool(torch.equal(got_rad.view(torch.int32), ref_rad.view(torch.int32)))
            b = bool(torch.equal(f_rgb, ref_rgb)) and bool(torch.equal(f_rad.view(torch.int32), ref_rad.view(torch.int32)))
            t = frame.timeouts()
            print("%-28s frame %d: shared-frame == 1-rank: %s | nccl gather == 1-rank: %s | spin timeouts %d | mean %.4f" % (
                name, rep, a, b, t, float(ref_rad.mean())), flush=True)
            ok = ok and a and b and t == 0 and float(ref_rad.abs().sum()) > 0
        dist.barrier()
    torch.cuda.synchronize()
    frame.close()


check("PATH cornell 200x120", 200, 120, abi.MODE_PATH, abi.SCENE_CORNELL, spp=8, max_depth=5)
check("PATH glass 173x90 (ragged)", 173, 90, abi.MODE_PATH, abi.SCENE_CORNELL_GLASS, spp=4, max_depth=8)
check("PATH 40x40 (tiles < ranks)", 40, 40, abi.MODE_PATH, abi.SCENE_CORNELL, spp=4, max_depth=3)
check("REF default 250x250", 250, 250, abi.MODE_REF, abi.SCENE_DEFAULT)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(flag, src=0)
dist.destroy_process_group()
if rank == 0:
    print("dist_check:", "OK" if ok else "FAILED")
sys.exit(0 if int(flag.item()) == 1 else 1)
