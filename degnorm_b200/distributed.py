"""
Multi-worker plumbing of the NMF-OA path: how genes are partitioned over workers (one process per GPU) and the one
small exchange the algorithm needs -- the sum of 3p+1 per-sample doubles between outer iterations
(reference: degnorm/nmf_mpi.py:603-629 gene scatter, :690-713 and :797-838 star gathers through rank 0, which
this replaces by an all-reduce; every rank then derives the same scale factors redundantly).

Communicator adapters give the engine one interface over
  * an mpi4py-style communicator (what degnorm_mpi passes: `.rank`, `.size`, lowercase pickled send/recv/allreduce),
  * a torch.distributed process group (NCCL on GPUs, gloo in CPU tests),
  * no communicator at all (one worker).
"""
import numpy as np


def partition_bounds(n_genes, size):
    """Contiguous gene blocks per worker, exactly utils.split_into_chunks(all_genes, n=size) (utils.py:176-192,
    used at nmf_mpi.py:605-606): chunk size ceil(n/size); trailing workers may get nothing."""
    size = max(1, int(size))
    cs = int(np.ceil(n_genes / float(size))) if n_genes else 0
    out = []
    for w in range(size):
        lo = min(w * cs, n_genes)
        out.append((lo, min(lo + cs, n_genes)))
    return out


def balanced_partition(cost, size):
    """Genes -> workers by estimated work: longest-processing-time-first greedy on `cost`
    (p x candidate columns per gene), each gene to the least loaded worker so far.  Returns one ascending index array
    per worker; results are scattered back by these indices, so gene order stays the caller's.  The reference's
    contiguous blocks (partition_bounds) follow the annotation order, where one chromosome's long genes can land on
    one worker; an outer iteration lasts as long as its slowest worker."""
    import heapq
    size = max(1, int(size))
    cost = np.asarray(cost, dtype=np.float64)
    order = np.argsort(-cost, kind="stable")
    heap = [(0.0, w) for w in range(size)]
    owner = np.empty(len(cost), dtype=np.int64)
    for g in order:
        load, w = heapq.heappop(heap)
        owner[g] = w
        heapq.heappush(heap, (load + cost[g], w))
    return [np.flatnonzero(owner == w) for w in range(size)]


def parse_cpulist(text):
    """'0-3,8,10-11' (the format of /sys/devices/system/node/nodeN/cpulist) -> sorted list of CPU numbers."""
    cpus = set()
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return sorted(cpus)


def pin_to_gpu_numa_node(device_index, min_cpus=4, sysfs="/sys"):
    """One process per GPU: keep this process (and the pinned staging memory it will first-touch) on the NUMA node
    the GPU hangs off, so that eight ranks packing and uploading at once do not cross the socket interconnect.
    Best effort: returns a dict saying what was done; leaves the affinity alone when the node is unknown (-1), the
    sysfs files are missing, or fewer than `min_cpus` of the node's CPUs are in this process's current affinity."""
    import os
    import torch
    info = dict(pinned=False)
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        info["pci"] = bdf
        node = int(open(os.path.join(sysfs, "bus/pci/devices", bdf, "numa_node")).read().strip())
        info["node"] = node
        if node < 0:
            return info
        cpus = parse_cpulist(open(os.path.join(sysfs, "devices/system/node/node%d/cpulist" % node)).read())
        mine = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        info["cpus"] = len(mine)
        if len(mine) < min_cpus:
            return info
        os.sched_setaffinity(0, mine)
        info["pinned"] = True
    except Exception as e:          # (no sysfs, no such device, not permitted ...)
        info["error"] = "%s: %s" % (type(e).__name__, e)
    return info


class SoloComm(object):
    rank, size = 0, 1

    def allreduce_(self, t):
        return t

    def send_obj(self, obj, dest, tag=0):
        raise RuntimeError("single worker")

    def recv_obj(self, source, tag=0):
        raise RuntimeError("single worker")

    def barrier(self):
        pass


class MPIComm(object):
    """mpi4py communicator (or anything with the same lowercase object API, e.g. a test double)."""

    def __init__(self, comm):
        self.comm = comm
        self.rank = comm.rank if hasattr(comm, "rank") else comm.Get_rank()
        self.size = comm.size if hasattr(comm, "size") else comm.Get_size()

    def allreduce_(self, t):
        # 3p+1 doubles: host round trip through the pickled allreduce (latency only)
        if self.size > 1:
            import torch
            host = t.detach().cpu().numpy()
            tot = self.comm.allreduce(host)          # default op: SUM
            t.copy_(torch.from_numpy(np.asarray(tot, dtype=np.float64)))
        return t

    def send_obj(self, obj, dest, tag=0):
        self.comm.send(obj, dest=dest, tag=tag)

    def recv_obj(self, source, tag=0):
        return self.comm.recv(source=source, tag=tag)

    def barrier(self):
        self.comm.Barrier()


class TorchComm(object):
    """torch.distributed process group; all_reduce runs on the tensor's device (NCCL) or through gloo."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)

    def allreduce_(self, t):
        if self.size > 1:
            self.dist.all_reduce(t, group=self.group)
        return t

    def _global(self, r):
        # send/recv_object_list address GLOBAL ranks; callers of this class speak group ranks
        return r if self.group is None else self.dist.get_global_rank(self.group, r)

    def send_obj(self, obj, dest, tag=0):
        # (torch has no message tags for objects: messages between a pair of ranks are matched in order)
        self.dist.send_object_list([obj], dst=self._global(dest), group=self.group)

    def recv_obj(self, source, tag=0):
        box = [None]
        self.dist.recv_object_list(box, src=self._global(source), group=self.group)
        return box[0]

    def barrier(self):
        self.dist.barrier(group=self.group)


def adapt(comm):
    if comm is None:
        return SoloComm()
    if isinstance(comm, (SoloComm, MPIComm, TorchComm)):
        return comm
    try:
        import torch.distributed as dist
        if isinstance(comm, dist.ProcessGroup):        # (it also has send / recv methods: test it first)
            return TorchComm(comm)
    except ImportError:
        pass
    if hasattr(comm, "send") and hasattr(comm, "recv"):
        return MPIComm(comm)
    raise TypeError("comm must be an mpi4py-style communicator, a torch.distributed process group or None")
