import queue
import threading

import numpy as np

_tl = threading.local()
_lock = threading.RLock()
_next_id = [1]


class Op:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return f"MPI.{self.name}"


SUM, MAX, MIN = Op("SUM"), Op("MAX"), Op("MIN")
_typedict = {c: c for c in "?bBhHiIlLqQefdgFDG"}
ANY_TAG = -1


def _reduce(vals, op):
    if op is SUM or op is None:
        acc = vals[0]
        for v in vals[1:]:
            acc = acc + v
        return acc
    if op is MAX:
        return np.maximum.reduce(vals) if isinstance(vals[0], np.ndarray) else max(vals)
    if op is MIN:
        return np.minimum.reduce(vals) if isinstance(vals[0], np.ndarray) else min(vals)
    raise NotImplementedError(op)


def _mem(a):
    """flat view of a contiguous array in MEMORY order (MPI moves raw bytes: an F-ordered
    buffer such as ``A.T.conj()`` travels in its own layout)"""
    a = np.asarray(a)
    if a.flags.c_contiguous:
        return a.reshape(-1)
    if a.flags.f_contiguous:
        return a.T.reshape(-1)
    raise ValueError("MPI buffers must be contiguous")


def _buf(spec):
    """mpi4py buffer spec: array or [array, count(s), (displs), type]"""
    return spec[0] if isinstance(spec, (list, tuple)) else spec


class Group:
    def __init__(self, ranks):
        self.ranks = list(ranks)

    def Incl(self, ranks):
        return Group([self.ranks[r] for r in ranks])


class Comm:
    """communicator over a fixed list of world ranks (threads)"""

    def __init__(self, members):
        self._members = list(members)
        with _lock:
            self._id = _next_id[0]
            _next_id[0] += 1
        n = len(self._members)
        self._barrier = threading.Barrier(n)
        self._slots = [None] * n
        self._children = {}
        self._queues = {}
        self._qlock = threading.Lock()

    # -- identity -------------------------------------------------------------
    def Get_rank(self):
        return self._members.index(_tl.rank)

    def Get_size(self):
        return len(self._members)

    rank = property(Get_rank)
    size = property(Get_size)

    def Get_group(self):
        return Group(self._members)

    def Barrier(self):
        self._barrier.wait()

    barrier = Barrier

    # -- object collectives ------------------------------------------------------
    def allgather(self, obj):
        r = self.Get_rank()
        self._slots[r] = obj
        self._barrier.wait()
        out = list(self._slots)
        self._barrier.wait()
        return out

    def allreduce(self, obj, op=SUM):
        return _reduce(self.allgather(obj), op)

    def bcast(self, obj, root=0):
        return self.allgather(obj)[root]

    # -- buffer collectives --------------------------------------------------------
    def Allgather(self, sendbuf, recvbuf):
        parts = self.allgather(np.array(_mem(_buf(sendbuf)), copy=True))
        _mem(_buf(recvbuf))[:] = np.concatenate(parts)

    def Allgatherv(self, sendbuf, recvspec):
        parts = self.allgather(np.array(_mem(_buf(sendbuf)), copy=True))
        rb, counts, displs = recvspec[0], recvspec[1], recvspec[2]
        flat = _mem(rb)
        for p, c, d in zip(parts, counts, displs):
            flat[d:d + c] = p[:c]

    def Allreduce(self, sendbuf, recvbuf, op=SUM):
        parts = self.allgather(np.array(_mem(_buf(sendbuf)), copy=True))
        _mem(_buf(recvbuf))[:] = _reduce(parts, op)

    def Bcast(self, buf, root=0):
        b = _buf(buf)
        val = self.allgather(np.array(_mem(b), copy=True) if self.Get_rank() == root else None)[root]
        if self.Get_rank() != root:
            _mem(b)[:] = val

    # -- point to point ---------------------------------------------------------------
    def _q(self, src, dst, tag):
        with self._qlock:
            return self._queues.setdefault((src, dst, tag), queue.Queue())

    def send(self, obj, dest, tag=0):
        self._q(self.Get_rank(), dest, tag).put(obj)

    def recv(self, source=0, tag=0):
        return self._q(source, self.Get_rank(), tag).get(timeout=60)

    def Send(self, buf, dest, tag=0):
        self._q(self.Get_rank(), dest, tag).put(np.array(_mem(_buf(buf)), copy=True))

    def Recv(self, buf, source=0, tag=0):
        b = _buf(buf)
        val = self._q(source, self.Get_rank(), tag).get(timeout=60)
        _mem(b)[:val.size] = val

    def Sendrecv(self, sendbuf, dest, sendtag=0, recvbuf=None, source=0, recvtag=0):
        self.Send(sendbuf, dest, sendtag)
        self.Recv(recvbuf, source, recvtag)

    def sendrecv(self, sendobj, dest, sendtag=0, recvbuf=None, source=0, recvtag=0):
        self.send(sendobj, dest, sendtag)
        return self.recv(source, recvtag)

    # -- communicator construction -------------------------------------------------------
    def _child(self, key, members):
        with _lock:
            c = self._children.get(key)
            if c is None:
                c = self._children[key] = Comm(members)
        return c

    def Split(self, color=0, key=0):
        seq = getattr(_tl, "seq", {})
        _tl.seq = seq
        n = seq.get(self._id, 0)
        seq[self._id] = n + 1
        info = self.allgather((color, key, _tl.rank))
        mine = sorted([i for i in info if i[0] == color], key=lambda t: (t[1], t[2]))
        return self._child((n, color), [m[2] for m in mine])

    def Create_group(self, group):
        seq = getattr(_tl, "seq", {})
        _tl.seq = seq
        n = seq.get(self._id, 0)
        seq[self._id] = n + 1
        return self._child((n, tuple(group.ranks)), group.ranks)


class _World:
    """MPI.COMM_WORLD: resolved per run (make_golden sets the active world)"""
    active = None

    def __getattr__(self, name):
        return getattr(_World.active, name)


COMM_WORLD = _World()


def run_world(size, fn, *args):
    """run fn(rank, *args) on `size` threads sharing a fresh COMM_WORLD; returns per-rank results"""
    _World.active = Comm(list(range(size)))
    results, errors = [None] * size, []

    def body(r):
        _tl.rank = r
        _tl.seq = {}
        try:
            results[r] = fn(r, *args)
        except BaseException as exc:  # noqa: BLE001
            errors.append((r, exc))
            try:
                _World.active._barrier.abort()
            except Exception:
                pass

    threads = [threading.Thread(target=body, args=(r,)) for r in range(size)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0][1]
    return results
