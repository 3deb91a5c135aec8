"""The reference's inference client/server contract (Client_Server.py:10-256) re-served by the B200 network.

Unmodified reference workers (numba MCTS in their own processes) keep talking to `Parallelized_Session` over the
same shared-memory mailboxes; the server process answers them with the CUDA trunk instead of onnxruntime:

  mailbox = float32 view of a SharedMemory block; [0] = flag (1.0 request pending, 0.0 idle / response ready);
  [1:1+n] = request payload (inputs flattened and concatenated in dict order) or response payload (outputs in
  `outputs_feed_info` order); the flag is written LAST on both sides (Client_Server.py:38-39,216-217).
  feed info: inputs `{"inputs": [[-1, H, W, C], np.float32]}`, outputs `{"policy": [-1, P], "value": [-1, 1]}`,
  exactly one -1 per shape, the batch dimension may be non-leading (transposition helpers, :122-149).

`start_server(inputs_feed_info, outputs_feed_info, shms, providers, file_path, per_process_wait_time)` keeps the
reference signature; `file_path` names a checkpoint written by `session.save_checkpoint` (the reference passes the
ONNX file there) and `providers` may carry the CUDA device ordinal as `[("GazExecutionProvider", {"device_id": 0})]`.
A `session_factory` keyword lets tests (and other evaluators) supply any object with the session duck type.
"""
import os
import time
from collections import deque
from multiprocessing.shared_memory import SharedMemory

import numpy as np


class Parallelized_Session:
    """Client side of one mailbox (Client_Server.py:10-55)."""

    def __init__(self, worker_id, shm, inputs_feed_info: dict, outputs_feed_info: dict):
        self.inputs_feed_info = inputs_feed_info
        self.outputs_feed_info = {name: [shape, int(np.prod(shape))] for name, shape in outputs_feed_info.items()}
        self.worker_id = worker_id
        self.shm = shm
        self.shared_arr = np.ndarray(shape=(shm.size // 4), dtype=np.float32, buffer=shm.buf)

    def run(self, output_names: list, input_feed: dict):
        if not list(input_feed.keys()) == list(self.inputs_feed_info.keys()):
            raise ValueError(
                f"input feed key's doesn't match in content and order to the input_feed_shape, "
                f"{self.inputs_feed_info.keys()}, {input_feed.keys()}")
        arr = self.shared_arr
        while arr[0] != 0.0:          # wait until the mailbox is idle
            pass
        data = np.concatenate([np.asarray(a).reshape(-1) for a in input_feed.values()], dtype=np.float32)
        arr[1:1 + len(data)] = data
        arr[0] = 1.0                  # flag last
        while arr[0] != 0.0:          # wait for the response
            pass
        outputs, start = [], 1
        for shape, n in self.outputs_feed_info.values():
            outputs.append(np.array(arr[start:start + n], copy=True).reshape(shape))
            start += n
        return outputs


def _batch_to_front(shape):
    """axes permutation that moves the batch axis of a [..., -1, ...] shape to the front, or None"""
    b = shape.index(-1)
    if b == 0:
        return None
    return [b] + [i for i in range(len(shape)) if i != b]


class Server:
    """Batching inference server (Client_Server.py:58-217) on top of a session-duck-typed evaluator."""

    def __init__(self, inputs_feed_info, outputs_feed_info, shared_memories, providers, file_path,
                 per_process_wait_time=0.001, session_factory=None):
        for shape, _ in inputs_feed_info.values():
            if not isinstance(shape, list):
                raise TypeError("The input's shape must be a list")
            if shape.count(-1) > 1:
                raise ValueError("There can only be 1 dimension for the batch! This means that only 1 (-1) batch dim can be included")
        for shape in outputs_feed_info.values():
            if not isinstance(shape, list):
                raise TypeError("The output's shape must be a list")
            if shape.count(-1) > 1:
                raise ValueError("There can only be 1 dimension for the batch! This means that only 1 (-1) batch dim can be included")
        # per input: (flat length, sample shape, dtype, permutation standard (batch-first) -> declared layout)
        self.inputs_cfg = {}
        for name, (shape, dtype) in inputs_feed_info.items():
            sample = [d for d in shape if d != -1]
            front = _batch_to_front(shape)
            to_declared = None if front is None else list(np.argsort(front))
            self.inputs_cfg[name] = (int(np.prod(sample)), sample, dtype, to_declared)
        # per output: permutation declared layout -> batch-first
        self.outputs_cfg = {name: _batch_to_front(shape) for name, shape in outputs_feed_info.items()}
        self.output_names = list(outputs_feed_info.keys())
        self.shms = shared_memories
        self.num_workers = len(self.shms)
        self.wait_time = per_process_wait_time * self.num_workers if self.num_workers > 1 else 0.0
        self.past_fills, self.past_wait_times = deque(), deque()
        self.batches = 0
        self.requests = 0
        if session_factory is not None:
            self.sess = session_factory()
        else:
            from .session import GazSession
            device = 0
            for pr in providers or []:
                if isinstance(pr, (tuple, list)) and len(pr) == 2 and isinstance(pr[1], dict):
                    device = int(pr[1].get("device_id", device))
            self.sess = GazSession(checkpoint=file_path, max_batch=max(8, self.num_workers), device=device)

    def compute_wait_time(self, alpha, avg_request_rate, min_wait=5e-6):
        """Client_Server.py:150-159: wait long enough that a fraction (1 - alpha) of the workers are in, clamp to
        [5 us, 6 ms]"""
        if avg_request_rate <= 0:
            return min_wait
        ratio = alpha * self.num_workers / avg_request_rate
        optimal_t = -np.log(ratio) / avg_request_rate if ratio < 1 else min_wait
        return max(min_wait, min(6e-3, optimal_t))

    def serve_once(self, arrs, stop=None):
        """one tick: gather pending requests, one forward pass, scatter the answers; returns the batch size"""
        active, seen = [], set()
        start = time.monotonic()
        while len(active) == 0 or (len(active) < len(arrs) and time.monotonic() - start <= self.wait_time):
            for i, a in enumerate(arrs):
                if i not in seen and a[0] == 1.0:
                    active.append(a)
                    seen.add(i)
            if len(active) == 0 and stop is not None and stop():
                return 0
        elapsed = time.monotonic() - start
        if self.num_workers > 1:
            self.past_fills.append(len(active))
            self.past_wait_times.append(elapsed)
            if len(self.past_wait_times) > 1000:
                self.past_wait_times.popleft()
                self.past_fills.popleft()
            total = sum(self.past_wait_times)
            self.wait_time = self.compute_wait_time(0.05, sum(self.past_fills) / total if total > 0 else 0.0)
        feed = {}
        off = 1
        for name, (n, sample, dtype, to_declared) in self.inputs_cfg.items():
            x = np.stack([a[off:off + n].reshape(sample) for a in active]).astype(dtype, copy=False)
            feed[name] = x if to_declared is None else x.transpose(to_declared)
            off += n
        outs = self.sess.run(self.output_names, input_feed=feed)
        outs = [o if self.outputs_cfg[nm] is None else o.transpose(self.outputs_cfg[nm])
                for nm, o in zip(self.output_names, outs)]
        for i, a in enumerate(active):
            flat = np.concatenate([np.asarray(o[i]).reshape(-1) for o in outs], dtype=np.float32)
            a[1:1 + len(flat)] = flat
            a[0] = 0.0                # flag last
        self.batches += 1
        self.requests += len(active)
        return len(active)

    def start(self, stop=None):
        arrs = [np.ndarray(shape=(shm.size // 4), dtype=np.float32, buffer=shm.buf) for shm in self.shms]
        while stop is None or not stop():
            self.serve_once(arrs, stop)


def start_server(inputs_feed_info, outputs_feed_info, shms, providers, file_path, per_process_wait_time=0.001,
                 session_factory=None):
    Server(inputs_feed_info, outputs_feed_info, shms, providers, file_path, per_process_wait_time,
           session_factory=session_factory).start()


def create_shared_memory(inputs_feed_info, outputs_feed_info, num_workers=os.cpu_count()):
    """one mailbox per worker, 4 * (max(1 + sum(inputs), 1 + sum(outputs)) + 1) bytes (Client_Server.py:233-243)"""
    n_in = 1 + sum(-int(np.prod(shape)) for shape, _ in inputs_feed_info.values())
    n_out = 1 + sum(-int(np.prod(shape)) for shape in outputs_feed_info.values())
    return [SharedMemory(create=True, size=4 * (int(max(n_in, n_out)) + 1)) for _ in range(num_workers)]


def convert_to_single_info(batched_info):
    """{name: [shape, dtype] | shape} with the -1 batch dimension replaced by 1 (Client_Server.py:246-256)"""
    out = {}
    for name, info in batched_info.items():
        shape = np.array(info[0] if isinstance(info[0], list) else info)
        shape[shape == -1] = 1
        out[name] = shape.tolist()
    return out
