"""Checkpoint directory naming -- same rule and function names as the reference
(/root/reference/utils/experiment_manager.py:8-40): one directory level `<flag>_<value>` per flag, so every
flag combination trains into its own directory."""
import fnmatch
import os

NOT_PATH = ['base_dir', 'display_test', 'test_set', 'data_glob', 'min_queue_examples', 'cuda_graph', 'out_dir', 'num_runs']


def make_checkpoint_path(base_path, FLAGS):
    """reference :8-16"""
    for k, v in FLAGS.flag_values_dict().items():
        if k not in NOT_PATH:
            base_path = base_path + '/' + k + '_' + str(v)
    return base_path


def list_all_checkpoints(base_path, FLAGS=None):
    """reference :18-27: every directory below base_path that holds a `checkpoint` state file."""
    paths = []
    for root, _dirnames, filenames in os.walk(base_path):
        for _ in fnmatch.filter(filenames, 'checkpoint'):
            paths.append(root)
    return paths


def set_flags_given_checkpoint_path(path, FLAGS):
    """reference :29-40: recover the flag values from the directory names.  Flag names contain underscores
    themselves (`nr_res_blocks_1`), so the longest registered flag name that prefixes a component wins (the
    reference's `param.split('_')` only works for underscore-free names)."""
    known = FLAGS.flag_values_dict()
    for comp in path.strip('/').split('/'):
        best = None
        for name in known:
            if comp.startswith(name + '_') and (best is None or len(name) > len(best)):
                best = name
        if best is None:
            continue
        raw = comp[len(best) + 1:]
        cur = known[best]
        if isinstance(cur, bool):
            val = raw in ('True', 'true', '1')
        else:
            val = type(cur)(raw)
        setattr(FLAGS, best, val)
