for d in 0 1 2 3; do echo "== BDDB200_RES_DEBUG=$d"; BDDB200_RES_DEBUG=$d python tools/gpu_debug_resident2.py 2>&1 | grep -E "K=2|K=7" | grep "first_call=False" ; done
