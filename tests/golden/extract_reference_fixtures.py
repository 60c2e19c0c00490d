"""Extracts the reference's own GFA test fixtures (src/test_gfa.rs, get_test_gfa_N) into
tests/golden/ref_test_gfa_N.gfa.  Run once in the build container (needs /root/reference);
the outputs are committed because /root/reference does not exist on the GPU box."""
import os
import re
import sys

src = open("/root/reference/src/test_gfa.rs").read()
out_dir = os.path.dirname(os.path.abspath(__file__))
for m in re.finditer(r"pub fn get_test_gfa_(\d+)\(\) -> Vec<String> \{(.*?)\n\}", src, re.S):
    n, body = m.group(1), m.group(2)
    lines = re.findall(r'"((?:[^"\\]|\\.)*)"', body.split("vec![", 1)[1])
    text = "\n".join(l.replace("\\t", "\t") for l in lines) + "\n"
    with open(os.path.join(out_dir, f"ref_test_gfa_{n}.gfa"), "w") as f:
        f.write(text)
    print(n, len(lines), file=sys.stderr)
