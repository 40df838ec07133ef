"""Second half of 0001 (run after `git apply next/0001-wire-lockedroom-playground.patch`): moves LockedRoom and
Playground from the oracle-only "next" tables to the product tables on the test side. Written as edits by pattern
rather than as a patch so that later additions to the next-tables do not break it."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IDS = ["MiniGrid-LockedRoom-v0", "MiniGrid-Playground-v0"]


def move_entries(path, src_dict, dst_dict):
    s = open(path).read()
    moved = []
    for env_id in IDS:
        m = re.search(r'^    "%s": .*\n' % re.escape(env_id), s[s.index(src_dict + " = {"):], re.M)
        assert m, (path, env_id)
        line = m.group(0)
        s = s.replace(line, "", 1)
        moved.append(line)
    i = s.index(dst_dict + " = {")
    j = s.index("\n}", i) + 1
    s = s[:j] + "".join(moved) + s[j:]
    open(path, "w").write(s)


def sub(path, old, new):
    s = open(path).read()
    assert s.count(old) == 1, (path, old)
    open(path, "w").write(s.replace(old, new))


os.chdir(ROOT)
move_entries("oracle/oracle.py", "NEXT_SPECS", "ENV_SPECS")
move_entries("oracle/gen_golden.py", "NEXT_ROLLOUTS", "ROLLOUTS")
kinds_old = '"lavagap", "distshift", "multiroom"'
kinds_new = '"lavagap", "distshift", "multiroom", "lockedroom", "playground"'
sub("tests/test_abi.py", "kinds = [\"empty\", \"doorkey\", \"crossing\", \"fourrooms\", " + kinds_old + "]",
    "kinds = [\"empty\", \"doorkey\", \"crossing\", \"fourrooms\", " + kinds_new + "]")
sub("tests/test_oracle_golden.py", "(\"empty\", \"doorkey\", \"crossing\", \"fourrooms\", " + kinds_old + ")",
    "(\"empty\", \"doorkey\", \"crossing\", \"fourrooms\", " + kinds_new + ")")
sub("tests/test_oracle_next.py", 'DEVICE_NEXT = ["MiniGrid-LockedRoom-v0", "MiniGrid-Playground-v0"]', "DEVICE_NEXT = []")
for env_id in IDS:
    subprocess.check_call(["git", "mv", f"tests/golden/next_rollout_{env_id}.npz", f"tests/golden/rollout_{env_id}.npz"])
print("moved", IDS, "to the product tables; now: python -m minigrid_b200._build && python -m pytest tests -m 'not gpu' -q")
