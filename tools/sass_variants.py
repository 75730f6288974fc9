import subprocess, re, collections, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sass = subprocess.check_output(["cuobjdump","-sass",os.path.join(ROOT,"zkstark_b200/libzkstark_b200.so")], text=True)
cur=None; mixes={}
for line in sass.splitlines():
    m=re.search(r"Function : (\S+)", line)
    if m: cur=m.group(1); continue
    if cur and 'tree_pass_kernelILi0E' in cur:
        m=re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m: mixes.setdefault(cur,collections.Counter())[m.group(1).split('.')[0]]+=1
for k,mix in sorted(mixes.items()):
    v=re.search(r'Li0ELi(\d+)E',k).group(1)
    print(v, sum(mix.values()), dict(mix.most_common(9)))
