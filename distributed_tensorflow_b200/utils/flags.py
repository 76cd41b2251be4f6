"""Command-line flag system (``dtf.app.flags``).

Capability parity: reference ``distributed_mnist.py:17-38``,
``example_between_graph.py:15-24``, ``example_in_graph.py:12-17`` define flags
with ``DEFINE_string/integer/float/bool`` and read them through a global
``FLAGS`` object that parses ``sys.argv`` lazily on first attribute access
(the scripts never call ``app.run``; ``main()`` is invoked directly).

Design: a single registry object; values are parsed once, on first access,
from ``sys.argv[1:]``.  Accepted spellings: ``--name=value``, ``--name value``,
``--bool_flag`` / ``--nobool_flag`` / ``--bool_flag=true|false|1|0``.
Unknown arguments are left untouched (kept in ``FLAGS.unparsed``), which keeps
``pytest``/``torchrun`` arguments from raising.
"""
from __future__ import annotations

import sys
from typing import Any, Callable, Dict, List, Optional

__all__ = [
    "FLAGS", "DEFINE_string", "DEFINE_integer", "DEFINE_float", "DEFINE_bool",
    "DEFINE_boolean", "flags", "FlagValues",
]


def _parse_bool(text: str) -> bool:
    t = str(text).strip().lower()
    if t in ("1", "true", "t", "yes", "y"):
        return True
    if t in ("0", "false", "f", "no", "n"):
        return False
    raise ValueError("not a boolean flag value: %r" % (text,))


class _Flag:
    __slots__ = ("name", "default", "help", "parser", "kind")

    def __init__(self, name: str, default: Any, help: str, parser: Callable[[str], Any], kind: str):
        self.name, self.default, self.help, self.parser, self.kind = name, default, help, parser, kind


class FlagValues:
    """Lazily-parsed flag container.  ``FLAGS.name`` triggers parsing once."""

    def __init__(self) -> None:
        object.__setattr__(self, "_defs", {})          # name -> _Flag
        object.__setattr__(self, "_values", {})        # name -> parsed value
        object.__setattr__(self, "_parsed", False)
        object.__setattr__(self, "unparsed", [])

    # -- definition -------------------------------------------------------
    def _define(self, flag: _Flag) -> None:
        defs: Dict[str, _Flag] = self._defs
        if flag.name in defs and defs[flag.name].kind != flag.kind:
            raise ValueError("flag %r redefined with a different type" % flag.name)
        defs[flag.name] = flag
        # A definition after parsing (module imported late) picks up argv too.
        if self._parsed:
            object.__setattr__(self, "_parsed", False)

    # -- parsing ----------------------------------------------------------
    def _parse(self, argv: Optional[List[str]] = None) -> List[str]:
        defs: Dict[str, _Flag] = self._defs
        values: Dict[str, Any] = {k: f.default for k, f in defs.items()}
        # explicit programmatic overrides survive a re-parse
        values.update({k: v for k, v in self._values.items() if k in getattr(self, "_overrides", ())})
        args = list(sys.argv[1:] if argv is None else argv)
        rest: List[str] = []
        i = 0
        while i < len(args):
            a = args[i]
            if a == "--":
                rest.extend(args[i:])
                break
            if not a.startswith("-") or a in ("-", "--"):
                rest.append(a)
                i += 1
                continue
            body = a.lstrip("-")
            name, eq, val = body.partition("=")
            name = name.replace("-", "_")
            if name in defs:
                f = defs[name]
                if f.kind == "bool":
                    if eq:
                        values[name] = _parse_bool(val)
                    elif i + 1 < len(args) and args[i + 1].lower() in ("true", "false", "0", "1"):
                        values[name] = _parse_bool(args[i + 1])
                        i += 1
                    else:
                        values[name] = True
                else:
                    if not eq:
                        if i + 1 >= len(args):
                            raise ValueError("flag --%s needs a value" % name)
                        val = args[i + 1]
                        i += 1
                    values[name] = f.parser(val)
            elif name.startswith("no") and name[2:] in defs and defs[name[2:]].kind == "bool" and not eq:
                values[name[2:]] = False
            else:
                rest.append(a)
            i += 1
        object.__setattr__(self, "_values", values)
        object.__setattr__(self, "_parsed", True)
        object.__setattr__(self, "unparsed", rest)
        return rest

    def __call__(self, argv: Optional[List[str]] = None) -> List[str]:
        """Explicit parse, absl style: ``FLAGS(sys.argv)`` (argv[0] is skipped)."""
        return self._parse(None if argv is None else list(argv[1:]))

    # -- access -----------------------------------------------------------
    def __getattr__(self, name: str) -> Any:
        if name.startswith("_"):
            raise AttributeError(name)
        if not self._parsed:
            self._parse()
        try:
            return self._values[name]
        except KeyError:
            raise AttributeError("unknown flag %r" % name) from None

    def __setattr__(self, name: str, value: Any) -> None:
        if not self._parsed:
            self._parse()
        self._values[name] = value
        ov = set(getattr(self, "_overrides", ()))
        ov.add(name)
        object.__setattr__(self, "_overrides", ov)

    def __contains__(self, name: str) -> bool:
        return name in self._defs

    def flag_values_dict(self) -> Dict[str, Any]:
        if not self._parsed:
            self._parse()
        return dict(self._values)

    def reset(self) -> None:
        """Forget all definitions and values (used by tests and examples)."""
        object.__setattr__(self, "_defs", {})
        object.__setattr__(self, "_values", {})
        object.__setattr__(self, "_parsed", False)
        object.__setattr__(self, "_overrides", set())
        object.__setattr__(self, "unparsed", [])

    def help_text(self) -> str:
        lines = []
        for f in self._defs.values():
            lines.append("  --%s (%s, default %r): %s" % (f.name, f.kind, f.default, f.help))
        return "\n".join(lines)


FLAGS = FlagValues()


def DEFINE_string(name: str, default: Optional[str], help: str = "") -> None:
    FLAGS._define(_Flag(name, default, help, str, "string"))


def DEFINE_integer(name: str, default: Optional[int], help: str = "") -> None:
    FLAGS._define(_Flag(name, default, help, int, "int"))


def DEFINE_float(name: str, default: Optional[float], help: str = "") -> None:
    FLAGS._define(_Flag(name, default, help, float, "float"))


def DEFINE_bool(name: str, default: Optional[bool], help: str = "") -> None:
    FLAGS._define(_Flag(name, default, help, _parse_bool, "bool"))


DEFINE_boolean = DEFINE_bool


class _FlagsModule:
    """``dtf.app.flags`` namespace object (mirrors ``tf.app.flags``)."""
    FLAGS = FLAGS
    DEFINE_string = staticmethod(DEFINE_string)
    DEFINE_integer = staticmethod(DEFINE_integer)
    DEFINE_float = staticmethod(DEFINE_float)
    DEFINE_bool = staticmethod(DEFINE_bool)
    DEFINE_boolean = staticmethod(DEFINE_bool)


flags = _FlagsModule()
