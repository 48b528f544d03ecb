#!/usr/bin/env python
"""Opcode mix of one kernel launch from an .ncu-rep source page: warp instructions executed per SASS opcode.
usage: ncu_opmix.py REPORT launch_index [pixels]"""
import csv, subprocess, sys, collections, re
def main(path, skip='0', pixels=None):
    out = subprocess.run(['ncu','-i',path,'--page','source','--csv','--launch-skip',skip,'--launch-count','1'],capture_output=True,text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print(rows[0][1][:140])
    hdr = rows[1]; ix = hdr.index('Instructions Executed'); isrc = hdr.index('Source'); ismp = hdr.index('# Samples')
    mix = collections.Counter(); smp = collections.Counter(); n = 0; static = collections.Counter()
    for r in rows[2:]:
        if len(r) <= ix: continue
        s = r[isrc].strip()
        s = re.sub(r'^@!?U?P\d+\s+', '', s)
        op = s.split()[0].rstrip(';') if s else '?'
        op = '.'.join(op.split('.')[:2]) if op.startswith(('LD','ST')) else op.split('.')[0]
        try: c = int(r[ix]); sm = int(r[ismp])
        except ValueError: continue
        mix[op] += c; smp[op] += sm; static[op] += 1; n += c
    px = float(pixels) if pixels else None
    print('total warp inst %d, static SASS %d%s' % (n, sum(static.values()), (', thread-inst/pixel %.1f' % (32*n/px)) if px else ''))
    for op, c in mix.most_common(28):
        print('  %-12s %12d %5.1f%%  static %5d  samples %6d%s' % (op, c, 100*c/n, static[op], smp[op], ('  %.2f/px' % (32*c/px)) if px else ''))
if __name__ == '__main__': main(*sys.argv[1:])
