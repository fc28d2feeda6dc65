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


def bench_shard(rank: int, bodies_per_gpu: int):
    """Weak-scaling shard of the synthetic workload: rank r owns global body indices [r*B, (r+1)*B)."""
    return rank * bodies_per_gpu, bodies_per_gpu
