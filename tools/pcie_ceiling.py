"""Host<->device link ceilings of the box, alone and with N GPUs copying AT THE SAME TIME (run under gpurun).

    python tools/pcie_ceiling.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/pcie_ceiling.py --json gpurun_out/pcie_ceiling_n8.json

One process per GPU (like bench.py).  Every rank owns 1 GiB pinned host buffers (input: plain pinned and
write-combined; output: plain pinned) and 1 GiB device buffers.  For each SUBSET of ranks (each GPU alone, pairs,
quads, all) the active ranks start together after a barrier and run, for ~REPS GiB each:
    h2d        cudaMemcpyAsync pinned -> device
    h2d_wc     same from write-combined pinned memory
    h2d_2d     cudaMemcpy2DAsync of every other 11,520-byte row (what csic_process_host issues for cfg4)
    d2h        device -> pinned
    duplex     h2d and d2h on two streams at once (what the e2e pipeline needs)
    duplex_2d  h2d_2d and d2h at once, bytes in the cfg4 ratio 3:2
The aggregate of a subset is sum(bytes) / max over its ranks of the elapsed time.  The result is the ceiling the
`e2e` figure of bench.py is quoted against (bench.py reads profiles/r2/pcie_ceiling*.json).
Nothing here is on the product path; cuda-python's runtime bindings are used for the copies."""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

from cuda.bindings import runtime as rt


def ck(res):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError(f"CUDA error {err}")
    return res[1] if len(res) == 2 else res[1:]


def host_memcpy_bandwidth(threads, mb=256, reps=4):
    """Aggregate memmove bandwidth of `threads` host threads (read + write bytes / s), GB/s."""
    n = mb << 20
    bufs = [(ctypes.create_string_buffer(n), ctypes.create_string_buffer(n)) for _ in range(threads)]
    for a, b in bufs:
        ctypes.memset(a, 1, n); ctypes.memset(b, 2, n)
    go = threading.Barrier(threads + 1)

    def work(i):
        a, b = bufs[i]
        go.wait()
        for _ in range(reps):
            ctypes.memmove(b, a, n)
        go.wait()
    th = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in th:
        t.start()
    go.wait(); t0 = time.perf_counter(); go.wait(); dt = time.perf_counter() - t0
    for t in th:
        t.join()
    return 2.0 * n * reps * threads / dt / 1e9


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout
    except Exception as e:  # noqa: BLE001
        return f"<{e}>"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default="")
    ap.add_argument("--gib", type=int, default=1)
    ap.add_argument("--reps", type=int, default=4)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    ck(rt.cudaSetDevice(local))
    n = args.gib << 30
    h_in = ck(rt.cudaHostAlloc(n, rt.cudaHostAllocDefault))
    h_wc = ck(rt.cudaHostAlloc(n, rt.cudaHostAllocWriteCombined))
    h_out = ck(rt.cudaHostAlloc(n, rt.cudaHostAllocDefault))
    ctypes.memset(h_in, 3, n); ctypes.memset(h_wc, 4, n); ctypes.memset(h_out, 5, n)
    d_in = ck(rt.cudaMalloc(n)); d_out = ck(rt.cudaMalloc(n))
    s1 = ck(rt.cudaStreamCreateWithFlags(rt.cudaStreamNonBlocking)); s2 = ck(rt.cudaStreamCreateWithFlags(rt.cudaStreamNonBlocking))
    H2D, D2H = rt.cudaMemcpyKind.cudaMemcpyHostToDevice, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost
    row = 11520
    rows2d = n // (2 * row)
    out_n = (rows2d * row) * 2 // 3          # cfg4: 2 output bytes per 3 shipped input bytes

    def k_h2d(): ck(rt.cudaMemcpyAsync(d_in, h_in, n, H2D, s1)); return n, 0
    def k_h2d_wc(): ck(rt.cudaMemcpyAsync(d_in, h_wc, n, H2D, s1)); return n, 0
    def k_h2d_2d(): ck(rt.cudaMemcpy2DAsync(d_in, row, h_in, 2 * row, row, rows2d, H2D, s1)); return rows2d * row, 0
    def k_d2h(): ck(rt.cudaMemcpyAsync(h_out, d_out, n, D2H, s2)); return 0, n
    def k_duplex(): k_h2d(); k_d2h(); return n, n
    def k_duplex_2d():
        k_h2d_2d(); ck(rt.cudaMemcpyAsync(h_out, d_out, out_n, D2H, s2)); return rows2d * row, out_n
    kinds = {"h2d": k_h2d, "h2d_wc": k_h2d_wc, "h2d_2d": k_h2d_2d, "d2h": k_d2h, "duplex": k_duplex, "duplex_2d": k_duplex_2d}

    def sync():
        ck(rt.cudaStreamSynchronize(s1)); ck(rt.cudaStreamSynchronize(s2))

    def barrier():
        if dist is not None:
            dist.barrier()

    def measure(active, kind):
        fn = kinds[kind]
        mine = rank in active
        if mine:
            fn(); sync()
        barrier()
        t0 = time.perf_counter(); up = dn = 0
        if mine:
            for _ in range(args.reps):
                a, b = fn(); up += a; dn += b
            sync()
        dt = time.perf_counter() - t0
        rec = [rank, mine, dt, up, dn]
        if dist is not None:
            allr = [None] * world
            dist.all_gather_object(allr, rec)
        else:
            allr = [rec]
        act = [r for r in allr if r[1]]
        tmax = max(r[2] for r in act)
        return {"h2d_gbs": round(sum(r[3] for r in act) / tmax / 1e9, 2), "d2h_gbs": round(sum(r[4] for r in act) / tmax / 1e9, 2),
                "per_rank_gbs": [round((r[3] + r[4]) / r[2] / 1e9, 2) for r in act]}

    subsets = [[i] for i in range(world)]
    if world >= 2:
        subsets += [[0, 1]]
    if world >= 4:
        subsets += [[0, 2], [0, 3], [0, 1, 2, 3]]
    if world >= 8:
        subsets += [[0, 4], [0, 7], [4, 5, 6, 7], [0, 2, 4, 6], [0, 1, 4, 5], list(range(8))]
    results = []
    for sub in subsets:
        for kind in kinds:
            if len(sub) == 1 and sub[0] != 0 and kind in ("h2d_wc", "h2d_2d"):
                continue
            r = measure(sub, kind)
            r.update(gpus=sub, kind=kind)
            results.append(r)
            if rank == 0:
                print(f"gpus={sub} {kind:10s} H2D {r['h2d_gbs']:7.1f} GB/s  D2H {r['d2h_gbs']:7.1f} GB/s  per-rank {r['per_rank_gbs']}", flush=True)
    info = None
    if rank == 0:
        host_bw = {t: round(host_memcpy_bandwidth(t), 1) for t in (1, 2, 4, 8, 16, 32) if t <= (os.cpu_count() or 1)}
        print("host memmove GB/s (read+write) by threads:", host_bw, flush=True)
        info = {"world": world, "bytes_per_copy": n, "reps": args.reps, "results": results, "host_memmove_gbs": host_bw,
                "cpus": os.cpu_count(), "topo": sh("nvidia-smi topo -m"), "lscpu": sh("lscpu | head -40"),
                "numa_nodes": sh("cat /sys/devices/system/node/online"),
                "gpu_pci": sh("nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv"),
                "gpu_numa": sh("for d in /sys/bus/pci/devices/*; do if [ \"$(cat $d/vendor)\" = 0x10de ]; then echo $d $(cat $d/numa_node) $(cat $d/class); fi; done"),
                "meminfo": sh("head -5 /proc/meminfo"), "virt": sh("systemd-detect-virt 2>/dev/null; cat /sys/class/dmi/id/product_name 2>/dev/null"),
                "iommu": sh("ls /sys/kernel/iommu_groups 2>/dev/null | wc -l; cat /proc/cmdline")}
        if args.json:
            os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
            json.dump(info, open(args.json, "w"), indent=1)
    barrier()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
