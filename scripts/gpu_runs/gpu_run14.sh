cd $GRAFT_REPO_ROOT
PYDEM_B200_TS_DEBUG=45 timeout 900 python scripts/sweep_ab.py 4096 sweep=worklist sweep=tile,tile=1 sweep=tile,tile=0 sweep=tile,tile=3 sweep=tile,tile=4 sweep=tile,tile=6 > gpurun_out/r2_ab14.log 2>&1; grep -E '^\{|rror|^cond' gpurun_out/r2_ab14.log | grep cond; grep "late visits\|\[ts\] kernel\|CTA-time" gpurun_out/r2_ab14.log | awk 'NR%96>=7 && NR%96<=9'
grep timeline gpurun_out/r2_ab14.log | sed -n 3p
