"""Host-side plumbing for the sharded engine: one process per GPU.  The control plane is a few dozen bytes per run
(the 128-byte NCCL id that bootstraps the cudaIpc handle exchange, barriers around timed regions, a max over ranks), so
it is a plain TCP star here — no torch, no NCCL: rank 0 listens on MASTER_ADDR:MASTER_PORT+`port_offset`, the other ranks
connect.  The launcher's environment (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT, as set by torchrun) is all it needs.
The data path of the sharded FastSLAM step runs over peer memory inside the kernels (no collective call per step)."""
import ctypes as C
import os
import socket
import struct
import time


def shard_bounds(n_global, rank, world):
    """contiguous block [lo, hi) of the particle index space owned by `rank`"""
    if n_global % world != 0:
        raise ValueError("particle count must divide evenly over the ranks")
    nl = n_global // world
    return rank * nl, (rank + 1) * nl


class TcpGroup:
    """rank 0 = hub.  Every operation is a gather to the hub followed by a scatter of the result (world <= 16)."""

    def __init__(self, rank=None, world=None, addr=None, port=None, port_offset=23, timeout=120.0):
        self.rank = int(os.environ.get("RANK", 0)) if rank is None else rank
        self.world = int(os.environ.get("WORLD_SIZE", 1)) if world is None else world
        addr = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
        port = (int(os.environ.get("MASTER_PORT", 29500)) + port_offset) if port is None else port
        self.peers = []
        self.sock = None
        if self.world == 1:
            return
        if self.rank == 0:
            srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((addr, port))
            srv.listen(self.world)
            srv.settimeout(timeout)
            conns = {}
            while len(conns) < self.world - 1:
                c, _ = srv.accept()
                c.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                r = struct.unpack("<i", self._recvn(c, 4))[0]
                conns[r] = c
            srv.close()
            self.peers = [conns[r] for r in range(1, self.world)]
        else:
            t0 = time.time()
            while True:
                try:
                    s = socket.create_connection((addr, port), timeout=5.0)
                    break
                except OSError:
                    if time.time() - t0 > timeout:
                        raise
                    time.sleep(0.05)
            s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            s.settimeout(timeout)
            s.sendall(struct.pack("<i", self.rank))
            self.sock = s

    @staticmethod
    def _recvn(s, n):
        buf = b""
        while len(buf) < n:
            chunk = s.recv(n - len(buf))
            if not chunk:
                raise ConnectionError("peer closed the control connection")
            buf += chunk
        return buf

    def _exchange(self, payload, combine):
        """every rank contributes `payload` (bytes, fixed length); every rank gets combine([payload_0 .. payload_{w-1}])"""
        if self.world == 1:
            return combine([payload])
        n = len(payload)
        if self.rank == 0:
            parts = [payload] + [self._recvn(c, n) for c in self.peers]
            out = combine(parts)
            for c in self.peers:
                c.sendall(struct.pack("<i", len(out)) + out)
            return out
        self.sock.sendall(payload)
        m = struct.unpack("<i", self._recvn(self.sock, 4))[0]
        return self._recvn(self.sock, m)

    def barrier(self):
        self._exchange(b"\0", lambda p: b"\0")

    def broadcast_bytes(self, data, n):
        """rank 0's `data` (n bytes) to everybody"""
        return self._exchange(bytes(data) if self.rank == 0 else b"\0" * n, lambda p: p[0])

    def max(self, x):
        out = self._exchange(struct.pack("<d", float(x)), lambda p: struct.pack("<d", max(struct.unpack("<d", q)[0] for q in p)))
        return struct.unpack("<d", out)[0]

    def close(self):
        for c in self.peers:
            c.close()
        if self.sock:
            self.sock.close()


def broadcast_unique_id(group, make_id):
    """rank 0 calls make_id() -> 128 bytes (an ncclUniqueId); every rank gets the same bytes back."""
    raw = b""
    if group.rank == 0:
        raw = bytes(make_id())
        if len(raw) != 128:
            raise ValueError("ncclUniqueId must be 128 bytes")
    return group.broadcast_bytes(raw, 128)


def nccl_unique_id():
    from .api import _check, load_library
    L = load_library()
    buf = C.create_string_buffer(128)
    _check(L, L.pfgpu_nccl_unique_id(buf))
    return buf.raw
