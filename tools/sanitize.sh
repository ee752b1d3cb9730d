#!/bin/bash
# compute-sanitizer (memcheck, racecheck, initcheck) over the render kernel on small
# fixtures: staged and unstaged (QR_B200_NOSTAGE=1) scene, 640 / 768 / 1024-thread shapes.
# usage (on the GPU box): tools/sanitize.sh <out.log>
out=${1:-gpurun_out/sanitizer.log}
: > $out
for tool in memcheck racecheck initcheck; do
  for fx in test14_full_a4 test17_full_a4 synth1k_a4; do
    for nostage in 0 1; do
      for shape in 2 3 7; do
        # racecheck is the slow one: one shape per staging mode is enough for it
        if [ $tool = racecheck ] && [ $shape != 3 ]; then continue; fi
        echo "== $tool $fx nostage=$nostage shape=$shape" >> $out
        QR_B200_NOSTAGE=$nostage QR_B200_SHAPE=$shape timeout 600 compute-sanitizer --tool $tool \
            --error-exitcode 7 python tools/prof_run.py $fx 1 > /tmp/san.txt 2>&1
        echo "rc=$?" >> $out
        grep -E "ERROR SUMMARY|RACECHECK SUMMARY|pixels != reference|Error|hazard" /tmp/san.txt | head -8 >> $out
      done
    done
  done
done
grep -c "rc=0" $out
