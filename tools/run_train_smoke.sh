# train/flow_train.py end to end on one GPU: 1001 steps (saves a checkpoint at step 1000), resume from it, then a run whose
# graph does not fit the checkpoint (filter_size 16) -> the reference's fall-back to the random init (flow_train.py:61-72)
cd steady-state-flow-with-neural-nets_b200/train
rm -rf /tmp/ckpt_test
timeout 300 python flow_train.py --batch_size=8 --max_steps=1001 --base_dir=/tmp/ckpt_test --data_glob=/nonexistent 2>&1 | tail -4
echo "---- resume"
timeout 200 python flow_train.py --batch_size=8 --max_steps=1102 --base_dir=/tmp/ckpt_test --data_glob=/nonexistent 2>&1 | tail -4
echo "---- checkpoint that does not fit"
d=$(ls -d /tmp/ckpt_test/*/ | head -1)
timeout 200 python - "$d" <<'PY'
import sys, glob, numpy as np
f = sorted(glob.glob(sys.argv[1] + "/**/*.npz", recursive=True))[-1]
ck = dict(np.load(f))
k = "resnet_1_0_conv_1_conv/weights"
ck[k] = np.zeros((3, 3, 1, 16), np.float32)          # wrong width
np.savez(f, **ck)
print("corrupted", f, k)
PY
timeout 200 python flow_train.py --batch_size=8 --max_steps=1102 --base_dir=/tmp/ckpt_test --data_glob=/nonexistent 2>&1 | tail -4
