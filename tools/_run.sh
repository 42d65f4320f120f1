python -m pytest tests -m gpu -x -q -k "full_size_properties" 2>&1 | tail -4
bash tools/gpu_session.sh r02 bench launches ncu 2>&1 | grep -v "^total\|^drwx\|^-rw" | cut -c1-200
