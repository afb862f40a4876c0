for r in 1 2 3 5; do echo "ratio $r"; BPPGPU_WAVE_RATIO=$r python tools/e2e_breakdown.py config2 2>&1 | grep -v "waves 1:\|waves 8\|waves 6"; done
