python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest.txt
V1="NDB_IVF_TC_PHASES=1"
V2="NDB_IVF_TC_PHASES=2"
V5="ONE_BATCH=1 NDB_IVF_TC_EXPERIMENT_KEEP_BOUNDS=1 NDB_IVF_TC_PHASES=1"
NDB_B200_LIB_PATH=$PWD/neurondb_b200/lib/libndb_b200_ctr.so python tools/c4_probe.py c4 "$V1" "$V2" > gpurun_out/probe_c4_ctr.txt 2>&1
python tools/c4_probe.py c4 "$V1" "$V2" "$V5" > gpurun_out/probe_c4.txt 2>&1
python tools/c4_probe.py c2 "$V1" "$V2" "$V5" > gpurun_out/probe_c2.txt 2>&1
