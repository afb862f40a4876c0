for r in 1 2 3 4; do echo "ratio $r"; BPPGPU_WAVE_RATIO=$r python tools/e2e_breakdown.py config2 2>&1 | grep -v "waves 1:\|device only"; done
