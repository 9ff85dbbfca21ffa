#!/bin/bash
# After tools/r02_profile.sh ran under gpurun: turn gpurun_out/r02/final_* into the committed summaries under profiles/
cd /root/repo; O=gpurun_out/r02
python tools/ncu_summary.py $O/final_trace_c2.ncu-rep "ncu --set full --clock-control none, shipped round-2 kernel, C2 (cover 800x600x128)" > profiles/r02_trace_c2_ncu.txt
python tools/ncu_summary.py $O/final_trace_c4.ncu-rep "ncu --set full --clock-control none, shipped round-2 kernel, 10,000-sphere scene 960x540x16 (C4M)" > profiles/r02_trace_c4m_ncu.txt
python tools/ncu_stages.py $O/final_trace_c2.ncu-rep > profiles/r02_trace_c2_stages.txt
python tools/ncu_stages.py $O/final_trace_c4.ncu-rep > profiles/r02_trace_c4m_stages.txt
rm -f profiles/kernel_profile.json
python tools/ncu_profile_json.py $O/final_trace_c2.ncu-rep C2 profiles/kernel_profile.json > /dev/null
python tools/ncu_profile_json.py $O/final_trace_c4.ncu-rep C4M profiles/kernel_profile.json > /dev/null
cp $O/final_launches.csv profiles/r02_launches.csv
cp $O/final_memcheck.log profiles/r02_memcheck.log; cp $O/final_racecheck.log profiles/r02_racecheck.log
cat $O/final_memcheck_stdout.log >> profiles/r02_memcheck.log; cat $O/final_racecheck_stdout.log >> profiles/r02_racecheck.log
cat profiles/r02_trace_c2_stages.txt; grep -E "duration|issue_active|thread_inst|registers|dram__bytes" profiles/r02_trace_c2_ncu.txt
