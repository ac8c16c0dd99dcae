"""bench.py runs the training-step record in child processes that form their OWN process group next to torchrun's.
Regression test (CPU, gloo, world size 2, under the real torch.distributed.run launcher): the children must rendezvous on
MASTER_PORT+1 although the parents carry torchrun's environment (TORCHELASTIC_USE_AGENT_STORE made them hang in round 2)."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PARENT = textwrap.dedent("""
    import os, subprocess, sys
    sys.path.insert(0, {root!r})
    import bench
    rank, ws = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    assert os.environ.get("TORCHELASTIC_USE_AGENT_STORE") == "True"      # the situation under test
    env = bench.child_group_env(rank, int(os.environ["LOCAL_RANK"]), ws)
    assert not any(k.startswith("TORCHELASTIC_") for k in env)
    child = (
        "import os, sys; sys.path.insert(0, %r); import torch, torch.distributed as dist\\n"
        "from pmt_learning_for_semantic_segmentation_and_disparity_b200 import sharding\\n"
        "w = sharding.init_world('gloo'); t = torch.tensor([float(w.rank + 1)]); dist.all_reduce(t)\\n"
        "print('CHILD', w.rank, w.world_size, float(t)); sharding.shutdown(w)" % {root!r})
    out = subprocess.run([sys.executable, "-c", child], env=env, capture_output=True, text=True, timeout=90)
    print(out.stdout.strip(), out.stderr.strip()[-300:], flush=True)
    assert out.returncode == 0
""")


def test_children_of_torchrun_ranks_form_their_own_group(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "parent.py"
    script.write_text(PARENT.format(root=ROOT))
    env = {k: v for k, v in os.environ.items() if not k.startswith("TORCHELASTIC_")}
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         env=env, capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    lines = sorted(l.strip() for l in res.stdout.splitlines() if l.startswith("CHILD"))
    assert lines == ["CHILD 0 2 3.0", "CHILD 1 2 3.0"], res.stdout[-2000:]
