"""SASS listing of one kernel (instruction text only, encodings stripped).
usage: python tools/sass_listing.py astr_b200/csrc/sweep2_k.o 'sweep2_kernelILi2ELi0ELi32ELi1ELb0ELi8E' > profiles/<name>.txt"""
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
names = [l.split()[-1] for l in subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.split("\n")
         if "Function :" in l and pat in l]
assert len(names) == 1, names
sass = subprocess.run(["cuobjdump", "-sass", "-fun", names[0], obj], capture_output=True, text=True).stdout
dem = subprocess.run(["c++filt", names[0]], capture_output=True, text=True).stdout.strip()
print(f"# cuobjdump -sass of {dem}\n# object {obj} (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a)")
for line in sass.split("\n"):
    m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        print(f"/*{m.group(1)}*/ {m.group(2).strip()} ;")
