#!/bin/bash
set -x
mkdir -p gpurun_out
AVL_TRACE=gpurun_out/r02_trace_rollout.json timeout 600 python tools/profile_step.py 150 > gpurun_out/r02_profile_frozen_h.txt 2>&1
python tools/trace_streams.py gpurun_out/r02_trace_rollout.json > gpurun_out/r02_trace_rollout_streams.txt 2>&1; cat gpurun_out/r02_trace_rollout_streams.txt | head -50
python - <<'PY' > gpurun_out/r02_trace_rollout_step_timeline.txt 2>&1
import json
tr=json.load(open('gpurun_out/r02_trace_rollout.json'))
ev=[e for e in tr['traceEvents'] if e.get('cat') in ('kernel','gpu_memcpy','gpu_memset') and 'dur' in e]
ev.sort(key=lambda e:e['ts'])
# find the audio_render kernels: one per step -> step boundaries
marks=[e['ts'] for e in ev if 'audio_render' in e['name']]
print('steps seen', len(marks))
if len(marks)>=6:
    a,b=marks[4],marks[5]
    print('step span us', b-a)
    for e in ev:
        if a<=e['ts']<b:
            print(f"{e['ts']-a:9.1f} +{e['dur']:7.1f} s{e['args'].get('stream')} {e['name'][:70]}")
PY
head -5 gpurun_out/r02_trace_rollout_step_timeline.txt
rm -f gpurun_out/r02_trace_rollout.json
bash tools/sanitizer.sh 2>&1 | tail -40
