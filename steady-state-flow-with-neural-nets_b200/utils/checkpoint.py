"""Checkpoints keyed by the TF variable names (what tf.train.Saver(tf.global_variables())
writes in /root/reference/train/flow_train.py:45,63-72,94-97).

One checkpoint = one `model.ckpt-<step>.npz` whose keys are exactly the names a TF-1.x run of the reference
would hold -- `<scope>/weights`, `<scope>/biases` in TF layouts (HWIO, [k,k,Cout,Cin], [Cin,F]), the Adam slots
`<var>/Adam` (m), `<var>/Adam_1` (v), `beta1_power`, `beta2_power` -- plus `global_step`; next to it the
`checkpoint` state file TF keeps (`model_checkpoint_path: "model.ckpt-<step>"`).  A dump of a real TF checkpoint
(`{n: reader.get_tensor(n) for n in reader.get_variable_to_shape_map()}` -> np.savez) therefore restores 1:1.
"""
from __future__ import annotations

import os
import re

import numpy as np
import torch

STATE_FILE = 'checkpoint'


def _np(t):
    return t.detach().to('cpu', torch.float32).numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def save(train_dir, step, variables, flat=None, beta1=0.9, beta2=0.999, prefix='model.ckpt'):
    """variables: name -> tensor (flow_architecture.global_variables()); flat: flow_backward.FlatParams or None."""
    os.makedirs(train_dir, exist_ok=True)
    out = {name: _np(v) for name, v in variables.items()}
    if flat is not None:
        m, v = _np(flat.m), _np(flat.v)
        for name, (off, k, shape) in flat.offsets.items():
            out[name + '/Adam'] = m[off:off + k].reshape(shape)
            out[name + '/Adam_1'] = v[off:off + k].reshape(shape)
        out['beta1_power'] = np.float32(beta1 ** (flat.step + 1))       # TF updates the powers after each apply
        out['beta2_power'] = np.float32(beta2 ** (flat.step + 1))
        out['adam_step'] = np.int64(flat.step)
    out['global_step'] = np.int64(step)
    base = '%s-%d' % (prefix, step)
    tmp = os.path.join(train_dir, base + '.tmp.npz')
    np.savez(tmp, **out)
    os.replace(tmp, os.path.join(train_dir, base + '.npz'))
    with open(os.path.join(train_dir, STATE_FILE), 'w') as f:
        f.write('model_checkpoint_path: "%s"\n' % base)
    return os.path.join(train_dir, base + '.npz')


def get_checkpoint_state(train_dir):
    """Path of the newest checkpoint named by the state file (tf.train.get_checkpoint_state), or None."""
    try:
        with open(os.path.join(train_dir, STATE_FILE)) as f:
            m = re.search(r'model_checkpoint_path:\s*"([^"]+)"', f.read())
    except OSError:
        return None
    if not m:
        return None
    path = os.path.join(train_dir, m.group(1) + '.npz')
    return path if os.path.exists(path) else None


def load(path):
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def is_slot(name):
    return name.endswith('/Adam') or name.endswith('/Adam_1') or name in ('beta1_power', 'beta2_power', 'global_step',
                                                                           'adam_step')


def restore(path_or_dict, store, flat=None, strict=True):
    """Copy a checkpoint into the variable store (in place for variables that already exist -- views into the flat
    training buffer stay valid -- and as the initial value for variables the lazy graph has not created yet).
    Raises ValueError on a shape mismatch or (strict) on names the store holds that the checkpoint lacks, which is
    when the reference falls back to a random init (flow_train.py:66-72).  Returns global_step."""
    ck = load(path_or_dict) if isinstance(path_or_dict, (str, os.PathLike)) else dict(path_or_dict)
    weights = {k: v for k, v in ck.items() if not is_slot(k)}
    for name, v in store.vars.items():
        if name not in weights:
            if strict:
                raise ValueError('checkpoint has no variable %r' % name)
            continue
        if tuple(v.shape) != tuple(weights[name].shape):
            raise ValueError('variable %s: checkpoint shape %s, graph shape %s' % (name, weights[name].shape, tuple(v.shape)))
    with torch.no_grad():
        for name, arr in weights.items():
            if name in store.vars:
                store.vars[name].copy_(torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)))
    store.preload({k: np.asarray(v, dtype=np.float32) for k, v in weights.items() if k not in store.vars})
    store.touch()
    if flat is not None and any(k.endswith('/Adam') for k in ck):
        with torch.no_grad():
            for name, (off, k, shape) in flat.offsets.items():
                if name + '/Adam' in ck:
                    flat.m[off:off + k].copy_(torch.from_numpy(np.ascontiguousarray(ck[name + '/Adam'], dtype=np.float32).reshape(-1)))
                    flat.v[off:off + k].copy_(torch.from_numpy(np.ascontiguousarray(ck[name + '/Adam_1'], dtype=np.float32).reshape(-1)))
        if 'adam_step' in ck:
            flat.step = int(ck['adam_step'])
        elif 'beta1_power' in ck:                        # a TF dump: t = log(beta1_power) / log(beta1) - 1
            flat.step = max(0, int(round(np.log(float(ck['beta1_power'])) / np.log(0.9))) - 1)
    return int(ck.get('global_step', 0))
