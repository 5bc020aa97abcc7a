#!/bin/bash
# Round-end evidence in one gpurun call (one GPU): full default bench line, the ncu launch list of one
# step with time + DRAM bytes per launch, and one `ncu --set full` capture of the longest launch of the
# top kernels.  Everything lands in gpurun_out/ with the given prefix; profiles/ gets the summaries.
#   bash tools/collect_profiles.sh r2final
P=${1:-r2final}
O=gpurun_out
mkdir -p $O
python bench.py > $O/${P}_bench.json 2> $O/${P}_bench.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $O/${P}_launches_c5_100.csv python bench.py --ncu --warmup 0 --steps 1 > $O/${P}_ncu_list.log 2>&1
python tools/ncu_top.py c5_pde_100 $O/${P}_c5 front_cb_kernel chol_panel_update_kernel wide_fwd_upd_kernel wide_bwd_upd_kernel assemble_M_kernel chol_diag_kernel > $O/${P}_ncu_top.log 2>&1
for k in front_cb_kernel chol_panel_update_kernel wide_fwd_upd_kernel wide_bwd_upd_kernel assemble_M_kernel chol_diag_kernel; do
  f=$O/${P}_c5_$k.ncu-rep
  [ -f $f ] && ncu -i $f --page raw --csv > $O/${P}_c5_$k.raw.csv 2>/dev/null
done
python tools/level_profile.py c5_pde_100 > $O/${P}_levels_c5_100.txt 2>&1
python tools/level_profile.py c3_sparse_qp_n200k > $O/${P}_levels_c3.txt 2>&1
tail -c 400 $O/${P}_bench.json
