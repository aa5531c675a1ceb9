for V in -1 0 1 2 3; do echo "red_levels=$V"; POET_MSDA_RED_LEVELS=$V timeout 100 python tools/kernel_micro.py x 2>&1 | grep "msda bwd"; done
