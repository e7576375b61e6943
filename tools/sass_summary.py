import subprocess, re, collections
out=subprocess.run(["cuobjdump","-sass","/root/repo/livecell-instance-segmentation_b200/csrc/liblcr.so"],capture_output=True,text=True).stdout
kern=None; stats=collections.OrderedDict()
for line in out.splitlines():
    m=re.match(r"\s*Function : (\S+)",line)
    if m:
        kern=subprocess.run(["c++filt",m.group(1)],capture_output=True,text=True).stdout.strip().split('(')[0].replace('void ','')
        stats[kern]=collections.Counter(); continue
    m=re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)",line)
    if m and kern:
        op=m.group(1); stats[kern]['total']+=1
        for key in ('UBLKCP','UTMALDG','UTMASTG','SYNCS','RED','ATOM','LDG','LDS','STS','STG','FFMA2','BAR','UCGABAR','LDGSTS','TCGEN','UTCBAR','CCTL'):
            if op.startswith(key): stats[kern][key]+=1
print("| kernel | SASS instr | UBLKCP (bulk copy engine) | UTMALDG/UTMASTG (tensor-map TMA) | SYNCS (mbarrier) | RED | ATOM(S) | LDG | LDS | STS | STG | FFMA2 | cluster barriers |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
tot=collections.Counter()
for k,c in stats.items():
    if 'lcr::' not in k: continue
    tot.update(c)
    print(f"| `{k}` | {c['total']} | {c['UBLKCP']} | {c['UTMALDG']+c['UTMASTG']} | {c['SYNCS']} | {c['RED']} | {c['ATOM']} | {c['LDG']} | {c['LDS']} | {c['STS']} | {c['STG']} | {c['FFMA2']} | {c['UCGABAR']} |")
print(f"| **all** | {tot['total']} | {tot['UBLKCP']} | {tot['UTMALDG']+tot['UTMASTG']} | {tot['SYNCS']} | {tot['RED']} | {tot['ATOM']} | {tot['LDG']} | {tot['LDS']} | {tot['STS']} | {tot['STG']} | {tot['FFMA2']} | {tot['UCGABAR']} |")
print(); print(f"tcgen05 / TMEM instructions: {tot['TCGEN']+tot['UTCBAR']} (expected: nothing on this path is a dense contraction).  CCTL: {tot['CCTL']}.")
