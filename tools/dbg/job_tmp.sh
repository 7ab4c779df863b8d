tools/gpu_ab.sh ab libgato_b200_mb3.so libgato_b200_mb4.so
