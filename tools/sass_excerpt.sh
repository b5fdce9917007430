#!/bin/bash
# SASS evidence for profiles/: mnemonic counts of every kernel of the default route + the scatter batch of k_warp_ws<256,4>
SO=vr-video-generator_b200/libvrsbs.so
OUT=${1:-profiles/r02_sass_excerpt.md}
{
echo "# SASS excerpt of libvrsbs.so (cuobjdump -sass, sm_100a), default route"
echo
echo "| kernel | instructions | UBLKCP (bulk copy, TMA engine) | SYNCS (mbarrier) | ATOMS | BAR | REDUX | IDP | UTMALDG / UTMASTG | UTC*MMA |"
echo "|---|---|---|---|---|---|---|---|---|---|"
for k in k_depth_passILb1E k_depth_pass_f32 k_build_tables k_warp_wsILi256ELi4ELb0E k_warp_wsILi512ELi8ELb0E k_warp_wsILi256ELi4ELb1E k_band_list k_word_list k_blur_bandILi2ELi5ELi4E k_blur_sepILi3ELi9ELi8E k_blur_commit k_depth_lowres_tiledILb1E; do
  cuobjdump -sass $SO | awk -v k="$k" '/Function :/{f=index($0,k)>0} f{print}' > /tmp/_k.sass
  n=$(grep -cE '^\s+/\*[0-9a-f]{4,5}\*/' /tmp/_k.sass)
  c() { grep -cE "$1" /tmp/_k.sass; }
  echo "| \`$k\` | $n | $(c UBLKCP) | $(c SYNCS) | $(c 'ATOMS') | $(c ' BAR\.') | $(c REDUX) | $(c 'IDP') | $(c 'UTMALDG|UTMASTG') | $(c 'UTC.*MMA') |"
done
echo
echo "The row loads and stores of \`k_warp_ws\` are the non-tensor bulk copy (\`cp.async.bulk\`, SASS \`UBLKCP\`), the right TMA form for"
echo "contiguous rows; there is no tensor-map copy and no tensor-core instruction anywhere (the path is not a contraction)."
echo
echo "## k_warp_ws<256,4>: one scatter batch of 4 segments without wrap (81 instructions, 8 ATOMS.MAX)"
echo
echo '```'
cuobjdump -sass $SO | awk '/Function : /{f=index($0,"k_warp_wsILi256ELi4ELb0E")>0} f{print}' > /tmp/_ws.sass
python3 - <<'PY'
import re
ins=[]
for l in open('/tmp/_ws.sass'):
    m=re.search(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);',l)
    if m: ins.append((int(m.group(1),16), m.group(2).strip()))
idx=[i for i,(a,t) in enumerate(ins) if 'ATOMS' in t]
cl=[];cur=[idx[0]]
for i in idx[1:]:
    if i-cur[-1]<30: cur.append(i)
    else: cl.append(cur);cur=[i]
cl.append(cur)
best=None
for c in cl:
    if len(c)!=8: continue
    s=c[0]
    while s>0 and not re.match(r'(@!?U?P\d+\s+)?(BRA|BSYNC|BSSY|BAR|WARPSYNC|EXIT)',ins[s-1][1]): s-=1
    n=c[-1]-s+1
    if best is None or n<best[0]: best=(n,s,c[-1])
for a,t in ins[best[1]:best[2]+1]: print(f"/*{a:04x}*/ {t}")
PY
echo '```'
} > $OUT
echo wrote $OUT
