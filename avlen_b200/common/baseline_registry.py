"""Trainer registry with the reference's decorator API (ss_baselines/common/baseline_registry.py)."""
from __future__ import annotations


class BaselineRegistry:
    _trainers: dict = {}

    @classmethod
    def register_trainer(cls, to_register=None, *, name=None):
        def wrap(klass):
            cls._trainers[name or klass.__name__] = klass
            return klass
        return wrap if to_register is None else wrap(to_register)

    @classmethod
    def get_trainer(cls, name):
        return cls._trainers.get(name)


baseline_registry = BaselineRegistry()
