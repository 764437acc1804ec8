run() { echo -n "$1 "; RTK_B200_LIB=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_call'], d['roofline_dpselect']['ms_per_call'])"; }
D=$PWD/video-retake_b200/lib/librtk_b200.so
run default $D
run small0 $PWD/build/ab/librtk_b_small0.so
run small0_score0 $PWD/build/ab/librtk_c_small0_score0.so
run score0 $PWD/build/ab/librtk_d_score0.so
run default $D
python tests/probes/ab_pdl_dpselect.py
