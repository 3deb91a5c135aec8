"""The shared-memory mailbox contract of Client_Server.py, served by our Server (CPU: a fake evaluator through
`session_factory`; GPU: the CUDA network in a real server process)."""
import multiprocessing as mp
import threading

import numpy as np
import pytest

from grok_alpha_zero_b200 import Client_Server as CS
from hash_eval import hash_eval


class FakeSession:
    def run(self, output_names, input_feed):
        x = np.asarray(input_feed["inputs"])
        assert x.ndim == 4 and x.dtype == np.float32
        pol = np.stack([hash_eval(s, 225)[0] for s in x])
        val = np.array([[hash_eval(s, 225)[1]] for s in x], dtype=np.float32)
        return [pol, val]


def test_feed_info_helpers_and_mailbox_size():
    inp = {"inputs": [[-1, 15, 15, 2], np.float32]}
    out = {"policy": [-1, 225], "value": [-1, 1]}
    shms = CS.create_shared_memory(inp, out, 3)
    try:
        assert len(shms) == 3 and all(s.size >= 4 * (451 + 1) for s in shms)
        assert CS.convert_to_single_info(inp) == {"inputs": [1, 15, 15, 2]}
        assert CS.convert_to_single_info(out) == {"policy": [1, 225], "value": [1, 1]}
    finally:
        for s in shms:
            s.close(); s.unlink()
    with pytest.raises(ValueError):
        CS.Server({"inputs": [[-1, -1, 2], np.float32]}, out, [], None, None, session_factory=FakeSession)
    with pytest.raises(TypeError):
        CS.Server({"inputs": [(-1, 3), np.float32]}, out, [], None, None, session_factory=FakeSession)


def test_mailbox_round_trips_match_direct_evaluation():
    inp = {"inputs": [[-1, 15, 15, 2], np.float32]}
    out = {"policy": [-1, 225], "value": [-1, 1]}
    n_workers = 4
    shms = CS.create_shared_memory(inp, out, n_workers)
    stop_flag = threading.Event()
    server = CS.Server(dict(inp), dict(out), shms, None, None, per_process_wait_time=1e-4, session_factory=FakeSession)
    th = threading.Thread(target=server.start, args=(stop_flag.is_set,), daemon=True)
    th.start()
    try:
        errors = []

        def worker(w):
            try:
                sess = CS.Parallelized_Session(w, shms[w], CS.convert_to_single_info(inp), CS.convert_to_single_info(out))
                rng = np.random.RandomState(w)
                for _ in range(25):
                    x = rng.randint(-1, 2, size=(1, 15, 15, 2)).astype(np.float32)
                    p, v = sess.run(["policy", "value"], {"inputs": x})
                    rp, rv = hash_eval(x[0], 225)
                    assert p.shape == (1, 225) and v.shape == (1, 1)
                    assert np.array_equal(p[0], rp) and v[0, 0] == rv
                with pytest.raises(ValueError):
                    sess.run(["policy", "value"], {"wrong": x})
            except Exception as ex:  # noqa: BLE001
                errors.append(ex)

        ws = [threading.Thread(target=worker, args=(w,)) for w in range(n_workers)]
        for w in ws:
            w.start()
        for w in ws:
            w.join(60)
        assert not errors, errors
        assert server.requests == n_workers * 25 and server.batches <= server.requests
    finally:
        stop_flag.set()
        th.join(5)
        for s in shms:
            s.close(); s.unlink()


def test_non_leading_batch_dimension():
    inp = {"inputs": [[3, 3, -1, 2], np.float32]}   # batch axis third (transposition helpers, Client_Server.py:122-149)
    out = {"policy": [9, -1], "value": [-1, 1]}
    seen = {}

    class S:
        def run(self, names, input_feed):
            x = input_feed["inputs"]
            seen["shape"] = x.shape
            b = x.shape[2]
            return [np.arange(9 * b, dtype=np.float32).reshape(9, b), np.ones((b, 1), np.float32)]

    shms = CS.create_shared_memory(inp, out, 1)
    try:
        server = CS.Server(inp, out, shms, None, None, session_factory=S)
        arr = np.ndarray(shape=(shms[0].size // 4), dtype=np.float32, buffer=shms[0].buf)
        arr[1:19] = np.arange(18)
        arr[0] = 1.0
        assert server.serve_once([arr]) == 1
        assert seen["shape"] == (3, 3, 1, 2) and arr[0] == 0.0
        assert np.array_equal(arr[1:10], np.arange(9, dtype=np.float32)) and arr[10] == 1.0
    finally:
        shms[0].close(); shms[0].unlink()


def _gpu_server(inp, out, names, ckpt):
    shms = [mp.shared_memory.SharedMemory(name=n) for n in names]
    CS.start_server(inp, out, shms, [("GazExecutionProvider", {"device_id": 0})], ckpt, 1e-4)


@pytest.mark.gpu
def test_reference_style_workers_against_the_cuda_server(tmp_path):
    """Self_Play.run_self_play's wiring (Self_Play.py:325-341): parent creates mailboxes, a server PROCESS loads the
    checkpoint, clients get network outputs within the bf16 tolerance of the fp32 oracle."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
    from net_oracle import NetOracle
    from grok_alpha_zero_b200 import netspec, session
    spec = netspec.build_spec("connect4", "softmax", num_blocks=2)
    W = netspec.init_weights(spec, seed=4)
    ckpt = str(tmp_path / "model.npz")
    session.save_checkpoint(ckpt, spec, W)
    inp = {"inputs": [[-1, 6, 7, 4], np.float32]}
    out = {"policy": [-1, 7], "value": [-1, 1]}
    shms = CS.create_shared_memory(inp, out, 2)
    ctx = mp.get_context("spawn")
    proc = ctx.Process(target=_gpu_server, args=(inp, out, [s.name for s in shms], ckpt), daemon=True)
    proc.start()
    try:
        ref = NetOracle(spec, W)
        rng = np.random.RandomState(0)
        for w in range(2):
            sess = CS.Parallelized_Session(w, shms[w], CS.convert_to_single_info(inp), CS.convert_to_single_info(out))
            for _ in range(5):
                x = rng.randint(-1, 2, size=(1, 6, 7, 4)).astype(np.float32)
                p, v = sess.run(["policy", "value"], {"inputs": x})
                o = ref.forward(x.astype(np.int8))
                assert np.abs(p - o["policy"].numpy()).max() < 2e-2 and abs(float(v[0, 0]) - float(o["value"][0])) < 1e-2
    finally:
        proc.terminate()
        for s in shms:
            s.close(); s.unlink()
