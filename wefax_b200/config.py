"""Settings loader with the reference's file format (config.py:4-14).

The reference opens ``config/config.json`` relative to the current working
directory at import time; the decode path reads only
``notch_filter_settings.{notch_filter_frequency, notch_filter_quality_factor}``
(wefax.py:63-64).  This loader looks for the same file in the same place (so a
deployment's existing config keeps working) and falls back to the defaults
shipped with the package.  The file is read when a ``Config`` is created, i.e.
once per ``Demodulator.process()``, not at import.
"""
from __future__ import annotations

import json
import os

_PACKAGE_DEFAULT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config", "config.json")


class Config:
    def __init__(self, path: str | None = None):
        self.settings: dict = {}
        self.path = path
        self.read_config_file()

    def read_config_file(self) -> None:
        candidates = [self.path] if self.path else [os.path.join("config", "config.json"), _PACKAGE_DEFAULT]
        for cand in candidates:
            if cand and os.path.isfile(cand):
                with open(cand) as fh:
                    for key, value in json.load(fh).items():
                        self.settings[key] = value
                self.path = cand
                break
        else:
            raise FileNotFoundError("config/config.json not found")
        if "notch_filter_settings" not in self.settings:
            with open(_PACKAGE_DEFAULT) as fh:
                self.settings["notch_filter_settings"] = json.load(fh)["notch_filter_settings"]
