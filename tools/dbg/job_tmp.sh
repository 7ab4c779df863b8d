for V in rtmb3 rtmb4; do echo $V; GATO_B200_LIB=gato_b200/lib/variants/libgato_b200_$V.so python tools/rt_vs_compiled.py 512 32 iiwa14 2>&1 | tail -1 | cut -c300-700; done
