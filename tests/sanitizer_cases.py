"""Driver of tests/run_sanitizers.sh: parity cases through a sanitizer build of the emulation library.
usage: python tests/sanitizer_cases.py <library.so> <seeds per k>"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import cases  # noqa: E402
from autocycler_b200 import api  # noqa: E402
from parity_common import check_case  # noqa: E402

lib = api.load_library(sys.argv[1])
n = 0
for k in (3, 5, 9, 31, 51, 65, 91, 127):
    for seed in range(int(sys.argv[2])):
        check_case(lib, cases.random_case(9000 * k + seed, k), k)
        n += 1
print("cases", n, "OK")
