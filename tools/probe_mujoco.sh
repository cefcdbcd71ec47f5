#!/bin/bash
# Probe a GPU lease for the reference's physics dependency (mujoco 3.5.0, pyproject.toml:33 of the reference).
# Writes a transcript to gpurun_out/mujoco_probe.txt; the result decides whether the dynamics parity can be pinned.
out=gpurun_out/mujoco_probe.txt
mkdir -p gpurun_out
{
  echo "# date: $(date -u +%FT%TZ)  host: $(hostname)  nproc: $(nproc)"
  echo "## python -c 'import mujoco'"; python -c "import mujoco; print(mujoco.__version__)" 2>&1 | tail -2
  echo "## python -c 'import mujoco_mjx / mujoco_warp / dm_control / gymnasium'"
  for m in mujoco_mjx mujoco.mjx mujoco_warp dm_control gymnasium warp brax; do python -c "import $m; print('$m ok')" 2>&1 | tail -1; done
  echo "## ls baseline/_ref"; ls baseline/_ref 2>&1 | head
  echo "## ls /opt/wheelhouse | grep -i mujoco"; ls /opt/wheelhouse 2>&1 | grep -i -E "mujoco|glfw|pyopengl|absl|etils" ; echo "(rc $?)"
  echo "## pip install --no-index --find-links /opt/wheelhouse mujoco==3.5.0 --target /tmp/mj"
  python -m pip install --no-index --find-links /opt/wheelhouse --target /tmp/mj "mujoco==3.5.0" 2>&1 | tail -4
  echo "## pip download mujoco==3.5.0 (index; expected to fail: no network)"
  timeout 60 python -m pip download --no-deps -d /tmp/mjdl "mujoco==3.5.0" 2>&1 | tail -4
  echo "## find / -iname '*mujoco*' (outside the repo)"
  find / -xdev \( -iname '*mujoco*' -o -iname 'libmujoco*' -o -iname 'mjmodel.h' \) -not -path '/proc/*' -not -path "$PWD/*" -not -path '/root/repo/*' 2>/dev/null | head -20
  echo "(end of find)"
  echo "## nvidia-smi -L"; nvidia-smi -L
  echo "## lscpu | head"; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket"
} > $out 2>&1
cat $out
