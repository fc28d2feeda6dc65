"""Request → GPU placement: independent units, no exchange.

The reference mints `internalReqID = x-request-id + "-" + uuid` (internal/extproc/server.go:191) and both ext_proc streams of a
request carry it; hashing THAT id keeps the router-stream parse and the upstream-stream translate on the same device."""

FNV_OFFSET = 0xcbf29ce484222325
FNV_PRIME = 0x100000001b3
MASK = (1 << 64) - 1


def hash64(request_id: bytes) -> int:
    """FNV-1a 64 — the same function the cgo shim would use (cheap, no allocation)."""
    h = FNV_OFFSET
    for b in request_id:
        h = ((h ^ b) * FNV_PRIME) & MASK
    return h


def device_for_request(request_id: bytes, n_devices: int) -> int:
    return hash64(request_id) % n_devices


def hash64_u64_array(ids):
    """FNV-1a 64 over the 8 little-endian bytes of every id of a numpy uint64 array (the vectorised form of hash64 for numeric
    request ids; bench.py --config 5 routes 10 M requests with it)."""
    import numpy as np
    ids = np.asarray(ids, dtype=np.uint64)
    h = np.full(ids.shape, FNV_OFFSET, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for k in range(8):
            h = (h ^ ((ids >> np.uint64(8 * k)) & np.uint64(0xFF))) * np.uint64(FNV_PRIME)
    return h


def devices_for_requests(ids, n_devices: int):
    import numpy as np
    return (hash64_u64_array(ids) % np.uint64(n_devices)).astype(np.int64)


def bench_shard(rank: int, bodies_per_gpu: int):
    """Weak-scaling shard of the synthetic workload: rank r owns global body indices [r*B, (r+1)*B)."""
    return rank * bodies_per_gpu, bodies_per_gpu
