"""Second half of 0004 (run after 0001 + its script and `git apply next/0004-wire-postfilter-kinds.patch`): moves the
step post-filter ids (GoToDoor, GoToObject, Fetch, PutNear, RedBlueDoors, Memory) from the oracle-only "next" tables
to the product tables on the test side. Pattern-based, like 0001-tests.py."""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
sys.path.insert(0, ROOT)
from oracle.oracle import NEXT_SPECS  # noqa: E402

IDS = [i for i in NEXT_SPECS if "Dynamic-Obstacles" not in i]
assert len(IDS) == 18, IDS


def move_entries(path, src_dict, dst_dict, ids):
    s = open(path).read()
    moved = []
    for env_id in ids:
        m = re.search(r'^    "%s": .*\n' % re.escape(env_id), s[s.index(src_dict + " = {"):], re.M)
        if not m:
            continue  # not every id has a fixture
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


move_entries("oracle/oracle.py", "NEXT_SPECS", "ENV_SPECS", IDS)
move_entries("oracle/gen_golden.py", "NEXT_ROLLOUTS", "ROLLOUTS", IDS)
old = '"lockedroom", "playground"'
new = '"lockedroom", "playground", "gotodoor", "fetch", "redbluedoors", "gotoobject", "putnear", "memory"'
sub("tests/test_abi.py", old + "]", new + "]")
sub("tests/test_oracle_golden.py", old + ")", new + ")")
sub("tests/test_abi.py", "L.mg_create(9, 8, 8,", "L.mg_create(99, 8, 8,")  # 9 is a kind now
s = open("tests/test_oracle_next.py").read()
s = s.replace('DEVICE_NEXT += [i for i in NEXT_SPECS if i not in DEVICE_NEXT and "Dynamic-Obstacles" not in i]', "DEVICE_NEXT += []")
open("tests/test_oracle_next.py", "w").write(s)
for env_id in IDS:
    for f in glob.glob(f"tests/golden/next_rollout_{env_id}.npz"):
        subprocess.check_call(["git", "mv", f, f.replace("next_rollout_", "rollout_")])
print("moved", len(IDS), "ids to the product tables")
