bash tools/gpu_session.sh r02b bench launches ncu 2>&1 | grep -v "^total\|^drwx\|^-rw" | cut -c1-300
