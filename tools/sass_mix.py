#!/usr/bin/env python
"""Static SASS opcode mix per kernel of a .so/.cubin: sass_mix.py FILE [name-substring]"""
import subprocess, sys, re, collections
out = subprocess.run(['cuobjdump','-sass',sys.argv[1]],capture_output=True,text=True).stdout
want = sys.argv[2] if len(sys.argv) > 2 else ''
name=None; mixes=collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m: name = m.group(1); mixes[name]=collections.Counter(); continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(.*?);', line)
    if m and name:
        s = re.sub(r'^@!?U?P\d+\s+','',m.group(1).strip())
        op = s.split()[0]
        parts = op.split('.')
        k = parts[0]
        if k in ('LDG','LDS','STS','STG','LDL','STL'): k = k + '.' + ''.join(p for p in parts[1:] if p in ('64','128'))
        mixes[name][k]+=1
for n,c in mixes.items():
    if want not in n: continue
    dem = subprocess.run(['cu++filt',n],capture_output=True,text=True).stdout.strip()
    tot=sum(c.values())
    print('%s\n  total %d: %s' % (dem[:150], tot, '  '.join('%s %d'%(k,v) for k,v in c.most_common(22))))
