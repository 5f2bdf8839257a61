cd steady-state-flow-with-neural-nets_b200/train && timeout 200 python flow_train.py --batch_size=16 --max_steps=301 --base_dir=/tmp/ckpt_test --data_glob=/nonexistent 2>&1 | tail -8
