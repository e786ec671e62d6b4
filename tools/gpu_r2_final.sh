python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -3 > gpurun_out/tests_final.txt; cat gpurun_out/tests_final.txt
for net in sscd vit; do ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lp_${net}_r2d.csv python tools/layer_profile.py run $net 256 > /dev/null 2>&1; python tools/layer_profile.py report gpurun_out/lp_${net}_r2d.csv $net 256 > gpurun_out/layers_${net}_r2d.txt; tail -1 gpurun_out/layers_${net}_r2d.txt; done
python bench.py --config c3 --other-modes "" > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c3.json')); print('c3', d['value'], d['e2e']['value'], d['images_embedded_per_s'], d['check']['indices_equal'], d['clocks'])"
