"""A small reader for the subset of gin-config the reference's inference path uses.

gin-config is not installed here; `InferenceModel.__init__` (msd/inference.py:83-111) only
needs: macros (`NAME = value`, `%NAME`), bindings (`scope/mod.Class.param = value`, incl. the
indented block form `mod.Class:\n  param = value`), configurable references (`@mod.fn()` /
`@scope/mod.Class()`), python literals, `import`/`from ... import` lines (ignored) and
`include '...'` (resolved against search roots).  That subset parses every file under the
reference's gin/models/diffusion, gin/tasks and gin/audio_codecs, and a T5X operative
`config.gin`.
"""

from __future__ import annotations

import ast
import os
import re
from typing import Any, Dict, List, Optional, Tuple


class ConfigurableRef:
  """`@scope/mod.name` (evaluate=False) or `@scope/mod.name()` (evaluate=True)."""

  def __init__(self, scope: str, name: str, evaluate: bool):
    self.scope, self.name, self.evaluate = scope, name, evaluate

  def __repr__(self):
    s = f'{self.scope}/' if self.scope else ''
    return f'@{s}{self.name}' + ('()' if self.evaluate else '')


class MacroRef:
  def __init__(self, name: str):
    self.name = name

  def __repr__(self):
    return f'%{self.name}'


_REF_RE = re.compile(r'@((?:[A-Za-z_][\w]*/)*)([A-Za-z_][\w.]*)(\(\))?')
_MACRO_RE = re.compile(r'%([A-Za-z_][\w.]*)')


class GinConfig:
  """Parsed bindings + macros."""

  def __init__(self):
    self.macros: Dict[str, Any] = {}
    # (scope, configurable_name, param) -> value ; configurable_name is the dotted tail
    self.bindings: Dict[Tuple[str, str, str], Any] = {}

  # ---- parsing -------------------------------------------------------------
  def parse(self, text: str, search_roots: Optional[List[str]] = None) -> 'GinConfig':
    lines = self._logical_lines(text)
    block: Optional[Tuple[str, str]] = None
    for raw, indented in lines:
      line = raw.strip()
      if not line:
        continue
      if line.startswith(('import ', 'from ')):
        block = None
        continue
      if line.startswith('include '):
        block = None
        path = ast.literal_eval(line[len('include '):].strip())
        self._include(path, search_roots or [])
        continue
      if indented and block is not None and '=' in line:
        param, value = line.split('=', 1)
        self.bindings[(block[0], block[1], param.strip())] = self._value(value.strip())
        continue
      block = None
      if line.endswith(':') and '=' not in line:
        scope, name = self._split_scope(line[:-1].strip())
        block = (scope, name)
        continue
      if '=' not in line:
        raise ValueError(f'gin_lite: cannot parse line: {raw!r}')
      lhs, value = line.split('=', 1)
      lhs = lhs.strip()
      scope, name = self._split_scope(lhs)
      if '.' not in name:
        self.macros[name if not scope else f'{scope}/{name}'] = self._value(value.strip())
      else:
        conf, param = name.rsplit('.', 1)
        self.bindings[(scope, conf, param)] = self._value(value.strip())
    return self

  def _include(self, path: str, roots: List[str]) -> None:
    for root in [''] + roots:
      cand = os.path.join(root, path)
      if os.path.exists(cand):
        with open(cand) as f:
          self.parse(f.read(), roots)
        return
    raise FileNotFoundError(f'gin_lite: include {path!r} not found under {roots}')

  @staticmethod
  def _split_scope(s: str) -> Tuple[str, str]:
    if '/' in s:
      scope, name = s.rsplit('/', 1)
      return scope, name
    return '', s

  @staticmethod
  def _logical_lines(text: str) -> List[Tuple[str, bool]]:
    """Strip comments, join bracket/backslash continuations; keep an 'indented' flag."""
    out: List[Tuple[str, bool]] = []
    buf, depth, indented = '', 0, False
    for raw in text.splitlines():
      line = GinConfig._strip_comment(raw)
      if not buf:
        if not line.strip():
          continue
        indented = line[:1] in (' ', '\t')
      buf = (buf + ' ' + line.strip()) if buf else line.rstrip()
      depth = GinConfig._depth(buf)
      if buf.endswith('\\'):
        buf = buf[:-1]
        continue
      if depth > 0 or buf.rstrip().endswith('='):
        continue
      out.append((buf, indented))
      buf = ''
    if buf.strip():
      out.append((buf, indented))
    return out

  @staticmethod
  def _strip_comment(line: str) -> str:
    q = None
    for i, ch in enumerate(line):
      if q:
        if ch == q:
          q = None
      elif ch in '\'"':
        q = ch
      elif ch == '#':
        return line[:i]
    return line

  @staticmethod
  def _depth(s: str) -> int:
    d, q = 0, None
    for ch in s:
      if q:
        if ch == q:
          q = None
      elif ch in '\'"':
        q = ch
      elif ch in '([{':
        d += 1
      elif ch in ')]}':
        d -= 1
    return d

  def _value(self, s: str) -> Any:
    """Python literal with @refs and %macros embedded."""
    holders: Dict[str, Any] = {}

    def ref_sub(m):
      key = f'__gin_ref_{len(holders)}__'
      holders[key] = ConfigurableRef(m.group(1).rstrip('/'), m.group(2), bool(m.group(3)))
      return repr(key)

    def mac_sub(m):
      key = f'__gin_ref_{len(holders)}__'
      holders[key] = MacroRef(m.group(1))
      return repr(key)

    t = _REF_RE.sub(ref_sub, s)
    t = _MACRO_RE.sub(mac_sub, t)
    val = ast.literal_eval(t)

    def restore(v):
      if isinstance(v, str) and v in holders:
        return holders[v]
      if isinstance(v, (list, tuple)):
        return type(v)(restore(x) for x in v)
      if isinstance(v, dict):
        return {restore(k): restore(x) for k, x in v.items()}
      return v

    return restore(val)

  # ---- queries ---------------------------------------------------------------
  def query_macro(self, name: str) -> Any:
    """`gin.query_parameter('%NAME')`, with nested %macros resolved."""
    name = name.lstrip('%')
    if name not in self.macros:
      raise KeyError(f'gin_lite: macro %{name} is not defined')
    return self.resolve(self.macros[name])

  def resolve(self, v: Any) -> Any:
    if isinstance(v, MacroRef):
      return self.query_macro(v.name)
    if isinstance(v, (list, tuple)):
      return type(v)(self.resolve(x) for x in v)
    if isinstance(v, dict):
      return {self.resolve(k): self.resolve(x) for k, x in v.items()}
    return v

  def bindings_for(self, configurable: str, scope: str = '') -> Dict[str, Any]:
    """Parameters bound to `configurable` (matched on dotted suffix), scope-aware:
    bindings of enclosing scopes apply, inner scopes override."""
    out: Dict[str, Any] = {}
    scopes = ['']
    if scope:
      parts = scope.split('/')
      scopes += ['/'.join(parts[:i + 1]) for i in range(len(parts))]
    for sc in scopes:
      for (bscope, conf, param), val in self.bindings.items():
        if bscope == sc and _suffix_match(conf, configurable):
          out[param] = val
    return out


def _suffix_match(bound: str, wanted: str) -> bool:
  b, w = bound.split('.'), wanted.split('.')
  n = min(len(b), len(w))
  return b[-n:] == w[-n:]


def parse_config(text: str, search_roots: Optional[List[str]] = None) -> GinConfig:
  return GinConfig().parse(text, search_roots)
