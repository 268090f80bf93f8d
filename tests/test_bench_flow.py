"""CPU: the control flow of bench.py's multi-GPU set-up (run_ours, N > 1) -- staged attempts (column panels -> plain p2p step ->
NCCL all-gather), the tuple it hands to the timing code, the fallback record -- executed with fakes in place of torch.cuda,
torch.distributed, the C ABI and ShardedCsr.  The block is taken from bench.py's source text between two marker lines, so a typo
in it (the kind that cost round 2 its N >= 4 runs) fails here instead of on an 8-GPU box."""
import os
import sys
import textwrap
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "bench.py")).read()
BEGIN, END = "        def make_local(r, c, arrays):", "    for _ in range(max(args.warmup, 3)):\n        step()"


class FakeOp:
    def prebuilt(self, *a):
        return lambda: None

    def close(self):
        pass

    def __call__(self, *a):
        pass


class FakeT:
    def __init__(self, v=1.0):
        self.v = v

    def __sub__(self, o):
        return FakeT(0.0)

    def __truediv__(self, o):
        return FakeT(0.0)

    def reshape(self, *a):
        return self

    def item(self):
        return self.v


def run_block(world, failing_attempts):
    class FakeSh:
        n = 0

        def __init__(self, off, col, val, rank, world_, make_local, exchange, overlap, row_weight):
            FakeSh.n += 1
            self.mode = (exchange, overlap)
            if FakeSh.n <= failing_attempts:
                raise TypeError(f"boom {FakeSh.n}")
            self.panels = overlap and world_ >= 4
            self.panel_ops, self.panel_calls = [1, 2, 3, 4], [lambda: None] * 4
            self.local_op, self.x_full, self.rows, self.nnz, self.x_block, self.cols_padded = FakeOp(), None, 10, 100, 5, 20
            self.off = self.col = self.val = None

        def new_x_shard(self, x):
            return "xs"

        def new_y_shard(self):
            return FakeT()

        def make_step(self, *a, **k):
            return lambda: None

        def describe_exchange(self):
            return f"fake {self.mode}"

    cuda = types.SimpleNamespace(synchronize=lambda: None, empty_cache=lambda: None)
    torch = types.SimpleNamespace(cuda=cuda, zeros_like=lambda t: FakeT(), linalg=types.SimpleNamespace(norm=lambda t: FakeT(1.0)))
    dist = types.SimpleNamespace(ReduceOp=types.SimpleNamespace(MAX=0), barrier=lambda: None, all_reduce=lambda t, op=None: None)

    class Operator(FakeOp):
        def __init__(self, *a, **k):
            pass
    cs = types.SimpleNamespace(Api=lambda name: "api", SpMVOperator=Operator)
    ns = dict(ShardedCsr=FakeSh, torch=torch, dist=dist, cs=cs, api=types.SimpleNamespace(set_option=lambda *a: None), os=os, sys=sys,
              args=types.SimpleNamespace(exchange="auto", row_weight=2.0), rank=0, world=world, csr_bytes=lambda *a: 123)
    a, b = SRC.index(BEGIN), SRC.index(END)
    body = "off, col, val, x = 1, 2, 3, 4\n" + textwrap.dedent(SRC[a:b]) + "\nreturn sh.mode, local, exchange, args._npanels, kernel_bytes, step, e2e_inner, local_call, xs, ys\n"
    exec("def f():\n" + textwrap.indent(body, "    "), ns)
    return ns["f"]()


def test_markers_exist_once():
    assert SRC.count(BEGIN) == 1 and SRC.count(END) == 1 and SRC.index(BEGIN) < SRC.index(END)


@pytest.mark.parametrize("world,fails,mode,npanels", [(8, 0, ("auto", True), 4), (8, 1, ("auto", False), 1), (8, 2, ("allgather", False), 1),
                                                      (4, 0, ("auto", True), 4), (2, 0, ("auto", False), 1), (2, 1, ("allgather", False), 1)])
def test_staged_set_up(world, fails, mode, npanels, capsys):
    got_mode, local, exchange, n, kbytes, step, e2e_inner, local_call, xs, ys = run_block(world, fails)
    assert got_mode == mode and n == npanels and kbytes == 123 and exchange == f"fake {mode}"
    assert callable(step) and callable(e2e_inner) and callable(local_call) and xs == "xs"
    assert local["max_rel_diff_vs_cusparse_over_ranks"] == 0.0
    assert len(local.get("fallback_from", [])) == fails             # every abandoned attempt is named in the JSON line
    if fails:
        assert "falling back" in capsys.readouterr().err


def test_the_last_attempt_raises():
    with pytest.raises(TypeError):
        run_block(8, 3)
    with pytest.raises(TypeError):
        run_block(2, 2)
