"""CPU baseline of the hot path (TEST / BENCH INFRASTRUCTURE ONLY - never imported by the package).

Mirrors how the reference runs self-play on a CPU box (Self_Play.py:346-363 + Client_Server.py:162-217):
`n_workers` workers, one game each, every simulation's leaf goes to ONE shared inference server that
batches whatever requests are pending (<= n_workers) into a single forward pass.  Here the workers are
threads driving the C oracle (oracle/mcts_oracle.c, released GIL) and the server is the fp32 PyTorch
restatement of the Keras network (oracle/net_oracle.py) standing in for onnxruntime-CPU, which the image
does not have.  kind = "port" in bench.py's cpu_baseline.
"""
import os
import sys
import threading
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import oracle as orc  # noqa: E402
from net_oracle import NetOracle  # noqa: E402


class BatchServer:
    """Client_Server.Server restated with threads: collect pending requests, one forward per tick."""

    def __init__(self, net, n_workers, wait_s=1e-3):
        self.net, self.n_workers, self.wait_s = net, n_workers, wait_s
        self.cv = threading.Condition()
        self.pending = {}
        self.results = {}
        self.active = 0
        self.stop = False
        self.batches = 0
        self.requests = 0
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def request(self, wid, state):
        with self.cv:
            self.pending[wid] = np.array(state, dtype=np.float32)
            self.cv.notify_all()
            while wid not in self.results:
                self.cv.wait()
            return self.results.pop(wid)

    def _loop(self):
        while True:
            with self.cv:
                while not self.pending and not self.stop:
                    self.cv.wait()
                if self.stop:
                    return
                # Client_Server.py:170-171: wait until every worker is in or wait_time elapsed
                t0 = time.perf_counter()
                while len(self.pending) < self.active and time.perf_counter() - t0 < self.wait_s:
                    self.cv.wait(self.wait_s)
                ids = list(self.pending.keys())
                batch = np.stack([self.pending.pop(i) for i in ids])
            out = self.net.forward(batch)
            pol = out["policy"].numpy()
            val = out["value"].numpy().reshape(-1)
            with self.cv:
                for k, i in enumerate(ids):
                    self.results[i] = (pol[k], float(val[k]))
                self.batches += 1
                self.requests += len(ids)
                self.cv.notify_all()

    def close(self):
        with self.cv:
            self.stop = True
            self.cv.notify_all()
        self.thread.join()


def run_sample(game, spec, weights, n_workers, sims_per_step, steps, warmup, c_puct_init=2.5, opening=None,
               torch_threads=None):
    """`steps` timed steps of `sims_per_step` PUCT simulations in each of `n_workers` games.
    Returns dict(sims_per_s, ms_per_step, sims, evals, seconds, mean_batch)."""
    if torch_threads:
        torch.set_num_threads(torch_threads)
    net = NetOracle(spec, weights)
    server = BatchServer(net, n_workers)
    trees = []
    for w in range(n_workers):
        def ev(st, w=w):
            return server.request(w, st)
        g = orc.OracleGame(game)
        if opening is not None and w % 3 == 0:
            g.do_action(opening)
        t = orc.OracleTree(game, False, evaluator=ev, c_puct_init=c_puct_init)
        trees.append((g, t))

    # Free-running workers, as the reference's worker PROCESSES are (Self_Play.py:346-363): every worker thread runs its own
    # simulations back to back and only meets the others inside the inference server's batch; threads are started once per
    # phase and joined once at its end (no per-step barrier).
    def phase(fn):
        server.active = n_workers
        th = [threading.Thread(target=fn, args=(i,)) for i in range(n_workers)]
        for x in th:
            x.start()
        for x in th:
            x.join()

    phase(lambda i: trees[i][1].new_root(trees[i][0]))
    if warmup > 0:
        phase(lambda i: trees[i][1].iterate(sims_per_step * warmup))
    s0 = sum(t.n_sims for _, t in trees)
    e0 = sum(t.n_evals for _, t in trees)
    b0, r0 = server.batches, server.requests
    t0 = time.perf_counter()
    phase(lambda i: trees[i][1].iterate(sims_per_step * steps))
    dt = time.perf_counter() - t0
    sims = sum(t.n_sims for _, t in trees) - s0
    evals = sum(t.n_evals for _, t in trees) - e0
    nb = max(1, server.batches - b0)
    mean_batch = (server.requests - r0) / nb
    server.close()
    return dict(sims_per_s=sims / dt, ms_per_step=dt / steps * 1e3, sims=sims, evals=evals, seconds=dt,
                mean_batch=mean_batch)
