"""Reader for the reference's scene-description language, used by the CLI when the reference package
itself is not importable (e.g. on a GPU box that only carries this repository).

Grammar (as accepted by the reference's ``scene_file.parse_scene``, scene_file.py:433-697):

    scene       := { float | material | sphere | plane | camera | point_light }
    float       := "float" IDENT "(" number ")"
    material    := "material" IDENT "(" brdf "," pigment ")"
    brdf        := ("diffuse" | "specular") "(" pigment ")"
    pigment     := "uniform" "(" color ")" | "checkered" "(" color "," color "," number ")"
                 | "image" "(" STRING ")"
    sphere      := "sphere" "(" IDENT "," transformation ")"          (plane: same)
    camera      := "camera" "(" ("perspective"|"orthogonal") "," transformation "," number "," number ")"
    point_light := "point_light" "(" vector "," color "," number ")"
    transformation := factor { "*" factor }
    factor      := "identity" | "translation" "(" vector ")" | "scaling" "(" vector ")"
                 | ("rotation_x"|"rotation_y"|"rotation_z") "(" number ")"
    vector := "[" number "," number "," number "]"      color := "<" number "," number "," number ">"
    number := LITERAL | IDENT (a float variable)        comments: "#" to end of line

This is a regular-expression tokenizer plus a recursive-descent parser written for this repository;
host-side, runs once per scene, not part of the hot path.
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional, Tuple

from . import scene as S
from .hdrimage import read_pfm_image


class GrammarError(Exception):
    def __init__(self, file_name: str, line: int, col: int, message: str):
        super().__init__(f"{file_name}:{line}:{col}: {message}")
        self.file_name, self.line_num, self.col_num, self.message = file_name, line, col, message


_TOKEN = re.compile(r"""
    (?P<ws>[ \t\r]+) | (?P<nl>\n) | (?P<comment>\#[^\n]*)
  | (?P<number>[+-]?(?:\d+\.?\d*|\.\d+)(?:[eE][+]?\d+)?)
  | (?P<ident>[A-Za-z_][A-Za-z_0-9]*)
  | (?P<string>"[^"\n]*")
  | (?P<symbol>[()<>\[\],*])
""", re.VERBOSE)


def _tokenize(text: str, file_name: str) -> List[Tuple[str, str, int, int]]:
    out, pos, line, line_start = [], 0, 1, 0
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            raise GrammarError(file_name, line, pos - line_start + 1, f"Invalid character {text[pos]!r}")
        kind = m.lastgroup
        if kind == "nl":
            line, line_start = line + 1, m.end()
        elif kind not in ("ws", "comment"):
            out.append((kind, m.group(), line, m.start() - line_start + 1))
        pos = m.end()
    out.append(("eof", "", line, pos - line_start + 1))
    return out


class Scene:
    def __init__(self):
        self.world = S.World()
        self.camera = None
        self.materials: Dict[str, S.Material] = {}
        self.float_variables: Dict[str, float] = {}


class _Parser:
    def __init__(self, text: str, file_name: str, variables: Dict[str, float]):
        self.toks, self.i, self.file_name = _tokenize(text, file_name), 0, file_name
        self.scene = Scene()
        self.scene.float_variables = dict(variables)
        self.overridden = set(variables)

    # -- token helpers
    def _peek(self):
        return self.toks[self.i]

    def _next(self):
        tok = self.toks[self.i]
        self.i += 1
        return tok

    def _fail(self, tok, message):
        raise GrammarError(self.file_name, tok[2], tok[3], message)

    def _symbol(self, sym: str):
        tok = self._next()
        if tok[0] != "symbol" or tok[1] != sym:
            self._fail(tok, f"got '{tok[1]}' instead of '{sym}'")

    def _ident(self) -> str:
        tok = self._next()
        if tok[0] != "ident":
            self._fail(tok, f"got '{tok[1]}' instead of an identifier")
        return tok[1]

    def _keyword(self, allowed) -> str:
        tok = self._next()
        if tok[0] != "ident" or tok[1] not in allowed:
            self._fail(tok, f"expected one of {', '.join(allowed)} instead of '{tok[1]}'")
        return tok[1]

    def _number(self) -> float:
        tok = self._next()
        if tok[0] == "number":
            return float(tok[1])
        if tok[0] == "ident":
            if tok[1] not in self.scene.float_variables:
                self._fail(tok, f"unknown variable '{tok[1]}'")
            return self.scene.float_variables[tok[1]]
        self._fail(tok, f"got '{tok[1]}' instead of a number")

    def _triple(self, open_sym: str, close_sym: str):
        self._symbol(open_sym)
        a = self._number()
        self._symbol(",")
        b = self._number()
        self._symbol(",")
        c = self._number()
        self._symbol(close_sym)
        return a, b, c

    # -- grammar
    def _pigment(self):
        kind = self._keyword(("uniform", "checkered", "image"))
        self._symbol("(")
        if kind == "uniform":
            pig = S.UniformPigment(S.Color(*self._triple("<", ">")))
        elif kind == "checkered":
            c1 = S.Color(*self._triple("<", ">"))
            self._symbol(",")
            c2 = S.Color(*self._triple("<", ">"))
            self._symbol(",")
            pig = S.CheckeredPigment(c1, c2, int(self._number()))
        else:
            tok = self._next()
            if tok[0] != "string":
                self._fail(tok, f"got '{tok[1]}' instead of a string")
            with open(tok[1][1:-1], "rb") as f:
                pig = S.ImagePigment(read_pfm_image(f))
        self._symbol(")")
        return pig

    def _transformation(self):
        result = S.Transformation()
        while True:
            kw = self._keyword(("identity", "translation", "rotation_x", "rotation_y", "rotation_z", "scaling"))
            if kw != "identity":
                self._symbol("(")
                if kw == "translation":
                    result = result * S.translation(S.Vec(*self._triple("[", "]")))
                elif kw == "scaling":
                    result = result * S.scaling(S.Vec(*self._triple("[", "]")))
                else:
                    result = result * getattr(S, kw)(self._number())
                self._symbol(")")
            nxt = self._peek()
            if nxt[0] == "symbol" and nxt[1] == "*":
                self._next()
                continue
            return result

    def _shape(self, cls):
        self._symbol("(")
        tok = self._peek()
        name = self._ident()
        if name not in self.scene.materials:
            self._fail(tok, f"unknown material {name}")
        self._symbol(",")
        t = self._transformation()
        self._symbol(")")
        return cls(transformation=t, material=self.scene.materials[name])

    def parse(self) -> Scene:
        sc = self.scene
        while True:
            tok = self._next()
            if tok[0] == "eof":
                return sc
            if tok[0] != "ident":
                self._fail(tok, f"expected a keyword instead of '{tok[1]}'")
            what = tok[1]
            if what == "float":
                name_tok = self._peek()
                name = self._ident()
                self._symbol("(")
                value = self._number()
                self._symbol(")")
                if name in sc.float_variables and name not in self.overridden:
                    self._fail(name_tok, f"variable «{name}» cannot be redefined")
                if name not in self.overridden:
                    sc.float_variables[name] = value
            elif what == "material":
                name = self._ident()
                self._symbol("(")
                kind = self._keyword(("diffuse", "specular"))
                self._symbol("(")
                pig = self._pigment()
                self._symbol(")")
                self._symbol(",")
                emitted = self._pigment()
                self._symbol(")")
                brdf = S.DiffuseBRDF(pig) if kind == "diffuse" else S.SpecularBRDF(pig)
                sc.materials[name] = S.Material(brdf, emitted)
            elif what == "sphere":
                sc.world.add_shape(self._shape(S.Sphere))
            elif what == "plane":
                sc.world.add_shape(self._shape(S.Plane))
            elif what == "camera":
                if sc.camera is not None:
                    self._fail(tok, "You cannot define more than one camera")
                self._symbol("(")
                kind = self._keyword(("perspective", "orthogonal"))
                self._symbol(",")
                t = self._transformation()
                self._symbol(",")
                aspect = self._number()
                self._symbol(",")
                distance = self._number()
                self._symbol(")")
                sc.camera = (S.PerspectiveCamera(distance, aspect, t) if kind == "perspective"
                             else S.OrthogonalCamera(aspect, t))
            elif what == "point_light":
                self._symbol("(")
                pos = self._triple("[", "]")
                self._symbol(",")
                color = S.Color(*self._triple("<", ">"))
                self._symbol(",")
                radius = self._number()
                self._symbol(")")
                sc.world.add_light(S.PointLight(S.Point(*pos), color, radius))
            else:
                self._fail(tok, f"Unexpected token {what}")


def parse_scene_text(text: str, variables: Optional[Dict[str, float]] = None, file_name: str = "<scene>") -> Scene:
    return _Parser(text, file_name, variables or {}).parse()
