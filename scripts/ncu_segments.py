"""Per-barrier-segment breakdown of one kernel from an ncu source-page CSV.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass --kernel-name regex:NAME > k.csv
       python scripts/ncu_segments.py k.csv
"""
import csv, sys, collections

def main(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    segs = []
    cur = dict(n=0, inst=0, samp=0, ops=collections.Counter(), st=collections.Counter(), wf=0, wfi=0)
    tot_inst = tot_samp = 0
    opsum = collections.Counter(); opsamp = collections.Counter()
    for r in rows[2:]:
        if len(r) < len(hdr): continue
        src = r[ci['Source']].strip()
        toks = src.split()
        op = toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '')
        inst = int(r[ci['Instructions Executed']] or 0)
        samp = int(r[ci['# Samples']] or 0)
        cur['n'] += 1; cur['inst'] += inst; cur['samp'] += samp
        cur['ops'][op] += inst
        cur['wf'] += int(r[ci['L1 Wavefronts Shared']] or 0)
        cur['wfi'] += int(r[ci['L1 Wavefronts Shared Ideal']] or 0)
        for s in stalls: cur['st'][s[6:]] += int(r[ci[s]] or 0)
        tot_inst += inst; tot_samp += samp
        opsum[op] += inst; opsamp[op] += samp
        if op.startswith('BAR'):
            segs.append(cur)
            cur = dict(n=0, inst=0, samp=0, ops=collections.Counter(), st=collections.Counter(), wf=0, wfi=0)
    segs.append(cur)
    print(f'total inst {tot_inst} samples {tot_samp}')
    for i, s in enumerate(segs):
        if s['inst'] == 0: continue
        f2 = sum(v for k, v in s['ops'].items() if k.startswith('FFMA2'))
        f1 = sum(v for k, v in s['ops'].items() if k.startswith('FFMA') and not k.startswith('FFMA2'))
        lds = sum(v for k, v in s['ops'].items() if k.startswith('LDS'))
        sts = sum(v for k, v in s['ops'].items() if k.startswith('STS'))
        top = ' '.join(f'{k}:{100*v//max(1,s["samp"])}' for k, v in s['st'].most_common(5))
        print(f'seg{i:2d} static {s["n"]:5d} inst {100*s["inst"]/tot_inst:5.1f}% samp {100*s["samp"]/tot_samp:5.1f}% '
              f'ffma2 {100*f2/s["inst"]:3.0f}% ffma {100*f1/s["inst"]:3.0f}% lds {100*lds/s["inst"]:3.0f}% sts {100*sts/s["inst"]:3.0f}% '
              f'wf {s["wf"]:9d} ideal {s["wfi"]:9d} | {top}')
    print('instruction mix')
    for k, v in opsum.most_common(28):
        print(f'{k:18s} {v:10d} {100*v/tot_inst:6.2f}%  samples {100*opsamp[k]/tot_samp:6.2f}%')

if __name__ == '__main__':
    main(sys.argv[1])
