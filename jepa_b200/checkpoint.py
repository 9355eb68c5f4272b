"""Asynchronous checkpoint writer.

The reference saves with a blocking torch.save of 5-10 GB on rank 0 every epoch (app/vjepa/train.py:307-324,340-346):
the step loop stalls for the device->host copy AND the serialisation AND the disk write.  Here the state is snapshotted
into pinned host buffers by asynchronous device->host copies enqueued on the compute stream (they are stream-ordered
BEFORE the next optimizer step, so the snapshot is consistent), and a background thread serialises it once the copies
have landed.  File format and keys are exactly the reference's (a dict of state_dicts written by torch.save)."""
import os
import threading

import torch


def _snapshot(obj, pool, memo):
    """Deep copy of a (nested) state dict with every CUDA tensor replaced by a pinned-host copy (async D2H); tensors that
    appear several times (the optimizer's shared device `step`) are copied once."""
    if torch.is_tensor(obj):
        hit = memo.get(id(obj))
        if hit is not None:
            return hit
        if obj.is_cuda:
            key = (tuple(obj.shape), obj.dtype)
            bufs = pool.setdefault(key, [])
            host = bufs.pop() if bufs else torch.empty(obj.shape, dtype=obj.dtype, pin_memory=True)
            host.copy_(obj.detach(), non_blocking=True)
        else:
            host = obj.detach().clone()
        memo[id(obj)] = host
        return host
    if isinstance(obj, dict):
        return type(obj)((k, _snapshot(v, pool, memo)) for k, v in obj.items())
    if isinstance(obj, (list, tuple)):
        return type(obj)(_snapshot(v, pool, memo) for v in obj)
    return obj


def _collect(obj, out):
    if torch.is_tensor(obj):
        if obj.is_pinned():
            out.append(obj)
    elif isinstance(obj, dict):
        for v in obj.values():
            _collect(v, out)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _collect(v, out)


class AsyncCheckpointer:
    def __init__(self):
        self._thread = None
        self._pool = {}          # recycled pinned buffers, keyed by (shape, dtype)
        self.error = None

    def save(self, save_dict, path):
        """Snapshot now (async copies on the current stream), write in the background.  Returns immediately."""
        self.wait()
        snap = _snapshot(save_dict, self._pool, {})
        done = torch.cuda.Event() if torch.cuda.is_available() else None
        if done is not None:
            done.record()

        def work():
            try:
                if done is not None:
                    done.synchronize()
                tmp = path + ".tmp"
                torch.save(snap, tmp)
                os.replace(tmp, path)
            except Exception as e:     # reported like the reference does (logged, training goes on)
                self.error = e
            finally:
                pinned, seen = [], set()
                _collect(snap, pinned)
                for t in pinned:
                    if id(t) not in seen:
                        seen.add(id(t))
                        self._pool.setdefault((tuple(t.shape), t.dtype), []).append(t)

        self._thread = threading.Thread(target=work, daemon=False)
        self._thread.start()

    def wait(self):
        if self._thread is not None:
            self._thread.join()
            self._thread = None
