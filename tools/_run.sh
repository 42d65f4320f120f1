# scratch command file for `gpurun -- 'bash tools/_run.sh'` (edited per experiment; the session script is tools/gpu_session.sh)
bash tools/gpu_session.sh scratch smoke tests bench
