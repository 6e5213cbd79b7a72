"""One-rank-per-GPU launch support: join the ranks of a torch.distributed job into a libb200kv peer
group (CUDA-IPC mapped NVLink peer memory, see csrc/group.h).

    torchrun --nproc-per-node 8 train.py
        import torch.distributed as dist, anand_mxnet_b200 as mx
        dist.init_process_group('nccl')            # or gloo
        mx.dist.init_peer_group()                  # once, before creating arrays / stores
        kv = mx.kv.create('device')                # kv.rank / kv.num_workers follow the job

torch.distributed is plumbing only: the library asks the host for ONE primitive, an all-gather of a
few bytes, used at group creation (IPC handles) and whenever a new call signature is planned
(operand offsets). Steady-state steps involve no host communication.
"""
import ctypes
import os

from .base import _LIB, check_call

_ALLGATHER_PROTO = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                    ctypes.c_void_p)
_state = {}


def make_allgather_callback(group=None):
    """A C callback that all-gathers `nbytes` over a (CPU / gloo) torch.distributed group."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)

    def _cb(send, recv, nbytes, _ctx):
        try:
            src = (ctypes.c_ubyte * nbytes).from_address(send)
            mine = torch.frombuffer(src, dtype=torch.uint8).clone()
            outs = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(world)]
            dist.all_gather(outs, mine, group=group)
            for r, o in enumerate(outs):
                ctypes.memmove(recv + r * nbytes, o.data_ptr(), nbytes)
            return 0
        except Exception as e:  # pragma: no cover - surfaced through the C error path
            print("b200kv all-gather callback failed:", e)
            return 1
    return _ALLGATHER_PROTO(_cb)


def nvls_wanted(world):
    """B200KV_NVLS = 1 / 0 forces the in-switch reduce on / off; unset or 'auto' uses it from 8
    ranks up (same rule as csrc/dense_group.cc)."""
    z = os.environ.get('B200KV_NVLS', 'auto')
    if z in ('', 'auto'):
        return world >= 8
    return z != '0'


def bind_to_gpu_numa_node(device_id):
    """Pin this rank's threads to the CPUs of the NUMA node its GPU hangs off, so that the pinned
    host buffers it allocates afterwards (first touch) and its PCIe traffic stay on that socket.
    With 8 ranks moving 2 x 102 MB per step each, cross-socket traffic is what limits the
    host-buffer path. Best effort: any missing piece (sysfs, torch attribute, container limits)
    leaves the affinity unchanged. B200KV_BIND_NUMA=0 disables it."""
    if os.environ.get('B200KV_BIND_NUMA', '1') in ('0', ''):
        return None
    try:
        import torch
        p = torch.cuda.get_device_properties(device_id)
        bdf = '%04x:%02x:%02x.0' % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open('/sys/bus/pci/devices/%s/numa_node' % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(','):
            a, _, b = part.partition('-')
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if len(allowed) >= 2:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:  # pragma: no cover - depends on the host
        pass
    return None


def init_peer_group(device_id=None, symmetric_memory=None):
    """Create the peer group for the calling torch.distributed job (idempotent).
    symmetric_memory=True (default when nvls_wanted()): take the arena from torch symmetric memory
    so that an NVSwitch multicast mapping is available to the kernels."""
    import torch.distributed as dist
    if _state.get('inited'):
        return
    assert dist.is_initialized(), "call torch.distributed.init_process_group first"
    rank, world = dist.get_rank(), dist.get_world_size()
    if device_id is None:
        device_id = int(os.environ.get('LOCAL_RANK', rank))
    _state['numa_node'] = bind_to_gpu_numa_node(device_id)
    # bootstrap traffic is a few hundred bytes on the host: always a gloo group
    cpu_group = dist.new_group(backend='gloo') if dist.get_backend() != 'gloo' else None
    cb = make_allgather_callback(cpu_group)
    _state.update(cb=cb, group=cpu_group, inited=True, rank=rank, world=world, device=device_id,
                  multicast=False)
    if symmetric_memory is None:
        symmetric_memory = nvls_wanted(world)
    if symmetric_memory and _init_with_symmetric_memory(rank, world, device_id, cb):
        return
    check_call(_LIB.B200KVGroupInit(ctypes.c_int(rank), ctypes.c_int(world), ctypes.c_int(device_id),
                                    cb, None))


def _init_with_symmetric_memory(rank, world, device_id, cb):
    """Arena from torch symmetric memory: torch allocates, peer-maps and (on NVSwitch fabrics)
    creates the MULTICAST mapping of all ranks' arenas; the library only receives the pointers.
    Needed for the NVLS mode of the fused kernel (B200KV_NVLS=1). Returns False when unavailable."""
    try:
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        mb = int(os.environ.get('B200KV_IPC_ARENA_MB', '6144'))
        dev = torch.device('cuda', device_id)
        arena = symm_mem.empty(mb << 20, dtype=torch.uint8, device=dev)
        hdl = symm_mem.rendezvous(arena, dist.group.WORLD)
        peers = [int(p) for p in hdl.buffer_ptrs]
        mc = int(hdl.multicast_ptr) if hdl.multicast_ptr else 0
        assert peers[rank] == arena.data_ptr()
    except Exception as e:  # pragma: no cover - depends on fabric / driver support
        print("b200kv: symmetric-memory arena unavailable (%s); using the CUDA-IPC arena" % (e,))
        return False
    peer_arr = (ctypes.c_void_p * world)(*peers)
    check_call(_LIB.B200KVGroupInitExternal(ctypes.c_int(rank), ctypes.c_int(world),
                                            ctypes.c_int(device_id), cb, None,
                                            ctypes.c_void_p(arena.data_ptr()),
                                            ctypes.c_size_t(arena.numel()), peer_arr,
                                            ctypes.c_void_p(mc or None)))
    _state.update(arena=arena, symm_handle=hdl, multicast=bool(mc))
    return True


def has_multicast():
    return bool(_state.get('multicast'))


def destroy_peer_group():
    if _state.get('inited'):
        check_call(_LIB.B200KVGroupDestroy())
        _state.clear()


def rank():
    return _state.get('rank', 0)


def world_size():
    return _state.get('world', 1)
