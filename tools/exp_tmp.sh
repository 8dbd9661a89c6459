for v in 0 1 2 3 4 5 6 7; do for mb in 4 6; do BCS_CP_V=$v BCS_CP_MB=$mb python tools/bench_stage.py springs 40; done; done
for g in 4 2; do for v in 0 4 7; do BCS_SPRING_G=$g BCS_CP_V=$v BCS_CP_MB=6 python tools/bench_stage.py springs 40; done; done
python tools/bench_stage.py springs 40
