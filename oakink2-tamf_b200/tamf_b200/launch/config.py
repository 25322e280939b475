"""Flag / config registry of the launchers.

The reference launchers are driven by `config_reg.ConfigRegistry` (thirdparty/config_reg @3252878, src/config_reg/reg.py:107
register, :220 parse) plus the run-bookkeeping flags of `dev_fn.upkeep.ckpt` (src/dev_fn/upkeep/ckpt.py:26-99).  That
package is a git submodule of the reference and is not shipped here, so this module implements the subset the two
sampling launchers use, with the same command-line surface:

  --cfg FILE.yml            repeatable; nested YAML keys become dotted entries, later files win
  --<prefix>.<key> VALUE    command line over config over default (ConfigEntrySource.COMMANDLINE_OVER_CONFIG)
  list[str] entries         colon separated  (ConfigEntryCommandlineSeqPattern.COLON_SEP)
  list[int] entries         comma separated  (COMMA_SEP)
  bool entries              --flag sets True  (ConfigEntryCommandlineBoolPattern.SET_TRUE)
  --exp_id / --commit       ckpt_path = <cwd>/common/<prog>/<exp_id>, log_file = <ckpt_path>/log.txt; without --commit
                            the run is a dry run: nothing is written (ckpt.py:108-120)
"""
from __future__ import annotations

import argparse
import os
import re
import sys
import time
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional

import yaml

UNSET = object()
_SPECIAL = re.compile(r"\?\(([^)]*)\)")


def _split_outside_special(text: str, sep: str) -> List[str]:
    """Split on `sep`, except inside a `?(...)` group (`?(file:./asset/split/test.txt)` is one element)."""
    out, cur, depth, i = [], "", 0, 0
    while i < len(text):
        if text.startswith("?(", i):
            depth += 1
            cur += "?("
            i += 2
            continue
        c = text[i]
        if c == ")" and depth:
            depth -= 1
        if c == sep and not depth:
            out.append(cur)
            cur = ""
        else:
            cur += c
        i += 1
    out.append(cur)
    return out


@dataclass
class Entry:
    key: str
    category: Any = str
    default: Any = UNSET
    required: bool = False
    seq: Optional[str] = None  # ":" or "," for list categories
    abspath: bool = False
    cmdline_only: bool = False
    callback: Optional[Callable[[Any, "Registry"], Any]] = None
    desc: str = ""
    value: Any = field(default=UNSET, repr=False)


class Registry:
    def __init__(self, prog: str):
        self.prog = prog
        self.entries: Dict[str, Entry] = {}
        self.timestamp = time.time()

    # -- registration -------------------------------------------------------------------------
    def register(self, key: str, prefix: Optional[str] = None, category: Any = str, default: Any = UNSET,
                 required: bool = False, seq: Optional[str] = None, abspath: bool = False, cmdline_only: bool = False,
                 callback=None, desc: str = "") -> None:
        full = f"{prefix}.{key}" if prefix else key
        if full in self.entries:
            raise KeyError(f"config entry registered twice: {full}")
        if category in (List[str], list) and seq is None:
            seq = ":"
        self.entries[full] = Entry(full, category, default, required, seq, abspath, cmdline_only, callback, desc)

    # -- parsing ------------------------------------------------------------------------------
    def _convert(self, e: Entry, raw: Any) -> Any:
        cat = e.category
        if cat is bool:
            if isinstance(raw, str):
                return raw.strip().lower() in ("1", "true", "yes", "on")
            return bool(raw)
        if cat in (int, float, str):
            return None if raw is None else cat(raw)
        if cat == List[int] or cat == List[str]:
            el = int if cat == List[int] else str
            if isinstance(raw, str):
                raw = [p for p in _split_outside_special(raw, e.seq or (":" if el is str else ",")) if p != ""]
            return [el(v) for v in raw]
        return raw

    @staticmethod
    def _flatten(d: dict, prefix: str = "") -> Dict[str, Any]:
        out = {}
        for k, v in (d or {}).items():
            full = f"{prefix}.{k}" if prefix else str(k)
            if isinstance(v, dict):
                out.update(Registry._flatten(v, full))
            else:
                out[full] = v
        return out

    def parse(self, argv: Optional[List[str]] = None) -> "Registry":
        ap = argparse.ArgumentParser(prog=self.prog)
        ap.add_argument("--cfg", action="append", default=[], help="YAML config file (repeatable, later files win)")
        for full, e in self.entries.items():
            if e.category is bool:
                ap.add_argument(f"--{full}", action="store_true", default=UNSET, help=e.desc)
            else:
                ap.add_argument(f"--{full}", type=str, default=UNSET, help=e.desc)
        ns = vars(ap.parse_args(argv))
        from_cfg: Dict[str, Any] = {}
        for path in ns["cfg"]:
            with open(path) as f:
                from_cfg.update(self._flatten(yaml.safe_load(f) or {}))
        unknown = sorted(k for k in from_cfg if k not in self.entries)
        if unknown:
            print(f"[{self.prog}] config keys without a registered entry are ignored: {unknown}", file=sys.stderr)
        for full, e in self.entries.items():
            raw = ns.get(full, UNSET)
            if raw is UNSET and not e.cmdline_only and full in from_cfg:
                raw = from_cfg[full]
            if raw is UNSET:
                raw = e.default
            e.value = UNSET if raw is UNSET else self._convert(e, raw)
        for full, e in self.entries.items():  # callbacks in registration order (dependencies are registered first)
            if e.callback is not None:
                e.value = e.callback(e.value, self)
            if e.abspath and isinstance(e.value, str):
                e.value = os.path.abspath(e.value)
            if e.required and (e.value is UNSET or e.value is None):
                ap.error(f"--{full} is required")
        return self

    # -- access -------------------------------------------------------------------------------
    def get(self, full: str) -> Any:
        v = self.entries[full].value
        return None if v is UNSET else v

    def select(self, prefix: str) -> Dict[str, Any]:
        """All entries under `prefix.` as a plain dict (ConfigRegistry.select)."""
        pre = prefix + "."
        res = {k[len(pre):]: (None if e.value is UNSET else e.value) for k, e in self.entries.items() if k.startswith(pre)}
        if not res:
            raise KeyError(prefix)
        return res


def expand_special(text: str, prog: str, ts: float) -> str:
    """`?(prog)`, `?(ts)`, `?(ts:date)`, `?(ts:full)` inside exp_id (ckpt.py:31-60); unknown commands expand to ''."""
    def sub(m):
        cmd = m.group(1)
        if cmd == "prog":
            return prog
        if cmd == "ts:date":
            return time.strftime("%Y_%m%d", time.localtime(ts))
        if cmd in ("ts", "ts:full"):
            return time.strftime("%Y_%m%d_%H%M_%S", time.localtime(ts))
        return ""
    return _SPECIAL.sub(sub, text)


def expand_process_range(items: List[str]) -> List[str]:
    """`?(file:PATH)` entries of data.process_range are replaced by the non-empty lines of PATH (the reference's split
    files, asset/split/*.txt); other entries pass through."""
    out: List[str] = []
    for it in items or []:
        m = _SPECIAL.fullmatch(it.strip())
        if m and m.group(1).startswith("file:"):
            with open(m.group(1)[5:]) as f:
                out.extend(ln.strip() for ln in f if ln.strip())
        else:
            out.append(it)
    return out


# ---- run bookkeeping (dev_fn/upkeep/ckpt.py) ---------------------------------------------------
def reg_ckpt(reg: Registry, exp_id_default: Optional[str] = None) -> None:
    def cb_exp_id(v, r: Registry):
        if v is UNSET or v is None:
            return f"{r.prog}__" + time.strftime("%Y_%m%d_%H%M_%S", time.localtime(r.timestamp))
        return expand_special(v, r.prog, r.timestamp)

    reg.register("exp_id", category=str, default=UNSET if exp_id_default is None else exp_id_default, callback=cb_exp_id)
    reg.register("ckpt_path", category=str, cmdline_only=True,
                 callback=lambda v, r: os.path.normpath(os.path.join(os.getcwd(), "common", r.prog, r.get("exp_id"))))
    reg.register("log_file", category=str, cmdline_only=True,
                 callback=lambda v, r: os.path.join(r.get("ckpt_path"), "log.txt"))
    reg.register("commit", category=bool, default=False, cmdline_only=True, desc="run in commit mode")


def ckpt_extract(reg: Registry) -> Dict[str, Any]:
    return {k: reg.get(k) for k in ("exp_id", "ckpt_path", "log_file", "commit")}


def _rotate(path: str) -> None:
    if not os.path.exists(path):
        return
    n = 1
    while os.path.exists(f"{path}.{n}"):
        n += 1
    os.replace(path, f"{path}.{n}")


def ckpt_setup(ckpt_cfg: Dict[str, Any], logger, rank: Optional[int] = None) -> None:
    """Commit mode creates <ckpt_path> and logs to <ckpt_path>/log.txt; otherwise dry run (ckpt.py:108-122)."""
    if rank:
        return
    if ckpt_cfg["commit"]:
        import logging
        os.makedirs(ckpt_cfg["ckpt_path"], exist_ok=True)
        fh = logging.FileHandler(ckpt_cfg["log_file"])
        fh.setFormatter(logging.Formatter("%(asctime)s %(name)s %(levelname)s %(message)s"))
        logging.getLogger().addHandler(fh)
        logger.info("commit mode: setup ckpt")
    else:
        logger.info("dry run mode")
    logger.info("cmd: %s", " ".join(sys.argv))


def ckpt_opt(ckpt_cfg: Dict[str, Any], rank: Optional[int] = None, **sections) -> None:
    """Commit mode dumps the resolved options to <ckpt_path>/opt.yml, rotating an existing file (ckpt.py:141-149)."""
    if rank or not ckpt_cfg["commit"]:
        return
    opt_file = os.path.join(ckpt_cfg["ckpt_path"], "opt.yml")
    _rotate(opt_file)
    with open(opt_file, "w") as f:
        yaml.safe_dump(sections, f, sort_keys=False)


# ---- model.* entries (launch/param/model.py:18-86) ---------------------------------------------
MODEL_DEFAULTS = dict(input_dim=99, obj_input_dim=9, hand_shape_dim=10, obj_embed_dim=768, latent_dim=256, ff_size=1024,
                      num_layers=8, num_heads=4, dropout=0.1, activation="gelu")


def reg_model_param(reg: Registry, prefix: str = "model") -> None:
    for k, v in MODEL_DEFAULTS.items():
        reg.register(k, prefix=prefix, category=type(v), default=v)


def reg_mano_param(reg: Registry, prefix: str = "mano", ws_dir: str = "") -> None:
    """mano.mano_path (launch/param/mano.py): directory holding MANO_RIGHT.pkl / MANO_LEFT.pkl."""
    reg.register("mano_path", prefix=prefix, category=str, default=os.path.join(ws_dir, "asset", "mano_v1_2"), abspath=True)
