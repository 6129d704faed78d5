"""Host<->device copy ceiling of one bench step, no kernels: every rank copies `mb` MB host-to-device and `mb` MB
device-to-host from/to pinned memory, concurrently on two streams, all ranks at the same time.  The e2e figure of bench.py
cannot beat this; bench.py runs the same probe and reports e2e.ceiling / e2e.frac_of_ceiling.

usage: python tools/pcie_probe.py [--mb 25.1] [--iters 20] [--pin-numa]        (torchrun --nproc-per-node N for N ranks)
prints one JSON line per rank 0: GB/s per direction per rank and summed, and the NUMA / affinity facts that explain them."""
import argparse
import json
import os
import time


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=24.9)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--pin-numa", action="store_true", help="bind this rank to the CPUs of its GPU's NUMA node before allocating")
    a = ap.parse_args()
    import torch
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    facts = {"cpu_affinity": len(os.sched_getaffinity(0)), "cpus": os.cpu_count()}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        facts["gpu_numa_node"] = pynvml.nvmlDeviceGetNumaNodeId(h) if hasattr(pynvml, "nvmlDeviceGetNumaNodeId") else None
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        facts["pci"] = bus if isinstance(bus, str) else bus.decode()
        node = "/sys/bus/pci/devices/%s/numa_node" % facts["pci"].lower()[-12:]
        if os.path.exists(node):
            facts["sysfs_numa_node"] = int(open(node).read())
        facts["pcie_gen_width"] = "gen%d x%d" % (pynvml.nvmlDeviceGetCurrPcieLinkGeneration(h), pynvml.nvmlDeviceGetCurrPcieLinkWidth(h))
    except Exception as e:
        facts["nvml"] = repr(e)[:100]
    if a.pin_numa and facts.get("sysfs_numa_node", -1) >= 0:
        lst = open("/sys/devices/system/node/node%d/cpulist" % facts["sysfs_numa_node"]).read().strip()
        cpus = set()
        for part in lst.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        os.sched_setaffinity(0, cpus)
        facts["pinned_to"] = lst
    n = int(a.mb * 1e6)
    h_up = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_dn = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_up = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_dn = torch.zeros(n, dtype=torch.uint8, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def once(up=True, dn=True):
        if up:
            with torch.cuda.stream(s_up):
                d_up.copy_(h_up, non_blocking=True)
        if dn:
            with torch.cuda.stream(s_dn):
                h_dn.copy_(d_dn, non_blocking=True)

    def run(up, dn):
        for _ in range(3):
            once(up, dn)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(a.iters):
            once(up, dn)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return n * a.iters / dt / 1e9
    both, up, dn = run(True, True), run(True, False), run(False, True)
    if rank == 0:
        print(json.dumps({"ranks": world, "mb_per_direction": a.mb, "GBs_per_rank_each_direction_concurrent": both,
                          "GBs_per_rank_h2d_alone": up, "GBs_per_rank_d2h_alone": dn, "GBs_all_ranks_each_direction_concurrent": both * world,
                          "c2_round_trip_ceiling_MPix_s": world * 3840 * 2160 / (24.9e6 / (both * 1e9)) / 1e6, "rank0": facts}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
