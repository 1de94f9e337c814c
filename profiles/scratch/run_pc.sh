export PYTHONPATH=$PWD
timeout 500 python profiles/prof_coach.py > gpurun_out/prof_coach.log 2>&1; echo rc=$?; grep -n "iteration:" gpurun_out/prof_coach.log
