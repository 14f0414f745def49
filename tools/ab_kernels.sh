for lib in "$@"; do echo "== $lib"; SSTTS_LIB=$PWD/$lib python tools/kernel_probe.py 2>&1 | tail -1; done
