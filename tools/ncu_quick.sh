#!/bin/bash
# ncu_quick.sh <report> <mangled-kernel-substring> [topN]: headline metrics + samples per CUDA source line of one capture
rep=$1; kern=$2; top=${3:-30}
python tools/ncu_summary.py $rep /tmp/ncu_quick.csv smsp__average_warps_issue_stalled 2>&1 | grep -v "^#" | awk -F, '{print $1, $3}' | grep -E "duration|issue_active|pipe_alu|pipe_lsu|pipe_fma.sum|wavefronts_mem_shared.sum.pct|inst_executed.sum|registers|stalled" | sed 's/smsp__average_warps_issue_stalled_//; s/_per_issue_active.ratio//'
d=$(mktemp -d); (cd $d; ncu -i $OLDPWD/$rep --page source --csv > sass.csv 2>/dev/null; cuobjdump -xelf k_front.sm_100a.cubin $OLDPWD/multimodal_biometric_fingerprints_palms_b200/libfpb200.so > /dev/null; nvdisasm -g -c k_front.sm_100a.cubin > dis.txt)
python tools/ncu_lines.py $d/sass.csv $d/dis.txt $kern $top
