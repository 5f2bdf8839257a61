"""Minimal stand-in for tf.app.flags (the reference defines its flags with it at import time:
model/flow_net.py:20-27, train/flow_train.py:15-24, test/flow_speed_test.py:28-37).  Same names,
same defaults; values can be overridden from the command line as --name=value or by assignment."""
from __future__ import annotations

import sys


class _Flags:
    def __init__(self):
        object.__setattr__(self, "_values", {})
        object.__setattr__(self, "_help", {})
        object.__setattr__(self, "_types", {})

    def _define(self, name, default, help_, typ):
        if name in self._values:        # the reference re-defines script flags in several files
            return
        self._values[name] = default
        self._help[name] = help_
        self._types[name] = typ
        for a in sys.argv[1:]:
            if a.startswith("--" + name + "="):
                self._values[name] = self._parse(a.split("=", 1)[1], typ)

    @staticmethod
    def _parse(s, typ):
        if typ is bool:
            return s.lower() in ("1", "true", "t", "yes", "y")
        return typ(s)

    def __getattr__(self, name):
        try:
            return self._values[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self._values[name] = value

    def flag_values_dict(self):
        return dict(self._values)


FLAGS = _Flags()


def DEFINE_string(name, default, help_=""):
    FLAGS._define(name, default, help_, str)


def DEFINE_integer(name, default, help_=""):
    FLAGS._define(name, default, help_, int)


def DEFINE_float(name, default, help_=""):
    FLAGS._define(name, default, help_, float)


def DEFINE_bool(name, default, help_=""):
    FLAGS._define(name, default, help_, bool)
