python -m pytest tests -m gpu -x -q -k "track or particle or stochastic" > gpurun_out/pytest_track.log 2>&1; tail -3 gpurun_out/pytest_track.log
tools/ab.sh 300
