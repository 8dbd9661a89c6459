"""Scene writers in the reference's compile-time header format.

The reference takes its scene from `src/config/*.hpp` - a client app overwrites those headers and the server is
recompiled (reference README.md:32).  These functions emit the same format from a `Scene`, so that a generated vein
(workloads.long_vein, make_cylinder_vein ...) or a custom blood-cell type can be handed to the UNMODIFIED reference
build (oracle/build_ref.sh <name> <dir>) as well as to libbcs:

  write_vein_definition        -> vein_definition.hpp        (veinPositions, veinIndices, VeinEndingCenters / Radii;
                                                               format of the reference's src/config/vein_definition.hpp:12-31776)
  write_blood_cell_presets     -> blood_cell_presets.hpp     (Springs / Vertices / Indices / Normals lists per type;
                                                               src/config/blood_cell_presets.hpp, meta_factory/blood_cells_def_type.hpp:36-49)
  write_blood_cells_definition -> blood_cells_definition.hpp (UserDefinedBloodCellList of BloodCellDef<...>)

Fixed-point convention of the reference's `mp_float<V, P>` / `Spring<a, b, L, P>`: value = float(V) / 10^(P-1)
(meta_factory/mp_helpers.hpp:23-35).  Every number is checked to survive the round trip bit-exactly; a value that
cannot be represented raises.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import numpy as np

from .scene import CellDef, Scene


def _fixed(value: float, precision: int) -> int:
    """integer V with float32(V) / float32(10^(precision-1)) == float32(value), as the reference evaluates it"""
    scale = np.float32(10 ** (precision - 1))
    v = int(round(float(value) * float(scale)))
    for cand in (v, v - 1, v + 1):
        if abs(cand) < 2 ** 31 and np.float32(cand) / scale == np.float32(value):
            return cand
    raise ValueError(f"{value!r} has no exact mp_float<V, {precision}> representation")


def _best_fixed(value: float, precisions: Sequence[int] = (7, 6, 5, 4, 8, 9)) -> "tuple[int, int]":
    for p in precisions:
        try:
            return _fixed(value, p), p
        except ValueError:
            continue
    raise ValueError(f"{value!r}: no exact fixed-point representation with precisions {tuple(precisions)}")


def write_vein_definition(path: str, scene: Scene) -> None:
    pos = np.asarray(scene.vein_pos, np.float32).reshape(-1, 3)
    idx = np.asarray(scene.vein_indices, np.uint32).reshape(-1)
    centers = np.asarray(scene.ending_centers, np.float32).reshape(-1, 3)
    radii = np.asarray(scene.ending_radii, np.float32).reshape(-1)
    out: List[str] = ["#pragma once", '#include "../meta_factory/mp_helpers.hpp"', '#include "../utilities/constexpr_vec.hpp"', "",
                      "#include <array>", "#include <boost/mp11/list.hpp>", "", "using namespace boost::mp11;", "", "// Vein positions", "",
                      f"inline constexpr std::array<cvec, {len(pos)}> veinPositions {{"]
    # %.9g round-trips every float32
    rows = [", ".join(f"static_cast<float>({c:.9g})" for c in p) for p in pos]
    out.append(",\n".join(rows))
    out += ["};", "", f"inline constexpr std::array<unsigned int, {len(idx)}> veinIndices {{"]
    out.append(",\n".join(", ".join(str(int(v)) for v in idx[k:k + 6]) for k in range(0, len(idx), 6)))
    out += ["};", "", "using VeinEndingCenters = mp_list<"]
    rows = []
    for c in centers:
        # one precision per vector (the reference's endings use 4)
        for p in (4, 5, 6, 7, 3, 2):
            try:
                rows.append("mp_float3<%d, %d, %d, %d>" % (*(_fixed(v, p) for v in c), p))
                break
            except ValueError:
                continue
        else:
            raise ValueError(f"ending centre {c} has no exact fixed-point representation")
    out.append(",\n".join(rows))
    out += [">;", "", "using VeinEndingRadii = mp_list<"]
    out.append(",\n".join("mp_float<%d, %d>" % _best_fixed(r, (4, 5, 6, 7, 3, 2)) for r in radii))
    out += [">;", ""]
    with open(path, "w") as f:
        f.write("\n".join(out))


def _preset_lists(name: str, d: CellDef) -> List[str]:
    out = [f"using {name}_Springs = mp_list<"]
    rows = []
    for (a, b), length in zip(np.asarray(d.springs).reshape(-1, 2), np.asarray(d.spring_lengths, np.float32)):
        v, p = _best_fixed(length)
        rows.append(f"\tSpring<{int(a)}, {int(b)}, {v}>" if p == 7 else f"\tSpring<{int(a)}, {int(b)}, {v}, {p}>")
    out.append(",\n".join(rows))
    out += [">;", "", f"using {name}_Vertices = mp_list<"]
    rows = []
    for vtx in np.asarray(d.vertices, np.float32).reshape(-1, 3):
        for p in (7, 6, 5, 4, 8):
            try:
                ints = [_fixed(c, p) for c in vtx]
                rows.append("\tmp_float3<%d, %d, %d>" % tuple(ints) if p == 7 else "\tmp_float3<%d, %d, %d, %d>" % (*ints, p))
                break
            except ValueError:
                continue
        else:
            raise ValueError(f"model vertex {vtx} has no exact fixed-point representation")
    out.append(",\n".join(rows))
    # render mesh of the type: not part of the physics path; a minimal valid triangle list keeps the renderer's
    # templates instantiable (indices into the model vertices, one normal per vertex)
    n = d.particles_in_cell
    tris = [(k, (k + 1) % n, (k + 2) % n) for k in range(max(1, n - 2))] if n >= 3 else [(0, 0, 0)]
    out += [">;", "", f"using {name}_Indices = mp_list<"]
    out.append(",\n".join("\t" + ", ".join(f"mp_int<{i}>" for i in t) for t in tris))
    out += [">;", "", f"using {name}_Normals = mp_list<"]
    out.append(",\n".join("\tmp_float3<0, 0, 1000000, 7>" for _ in range(n)))
    out += [">;", ""]
    return out


def indices_in_cell(d: CellDef) -> int:
    n = d.particles_in_cell
    return 3 * (max(1, n - 2) if n >= 3 else 1)


def write_blood_cell_presets(path: str, defs: Sequence[CellDef], names: Sequence[str]) -> None:
    out = ["#pragma once", "", '#include "../meta_factory/blood_cells_def_type.hpp"', "", "#include <boost/mp11/list.hpp>",
           "#include <boost/mp11/mpl_tuple.hpp>", "#include <glm/vec3.hpp>", "", "namespace preset", "{", "\tusing namespace boost::mp11;", ""]
    done = set()
    for d, name in zip(defs, names):
        if name in done:
            continue
        done.add(name)
        out += _preset_lists(name, d)
    out += ["}", ""]
    with open(path, "w") as f:
        f.write("\n".join(out))


def write_blood_cells_definition(path: str, defs: Sequence[CellDef], names: Sequence[str], colors: Sequence[int] = ()) -> None:
    out = ["#pragma once", "", '#include "blood_cell_presets.hpp"', '#include "../meta_factory/blood_cells_def_type.hpp"', "",
           "#include <boost/mp11/list.hpp>", "", "", "namespace", "{", "\tusing namespace boost::mp11;", "", "\tusing UserDefinedBloodCellList = mp_list<"]
    rows = []
    for k, (d, name) in enumerate(zip(defs, names)):
        color = colors[k] if k < len(colors) else 15720158
        rows.append(f"\tBloodCellDef<{d.count}, {d.particles_in_cell}, {indices_in_cell(d)}, {color},\n"
                    f"\t\tpreset::{name}_Springs,\n\t\tpreset::{name}_Vertices,\n\t\tpreset::{name}_Indices,\n\t\tpreset::{name}_Normals>")
    out.append(",\n\n".join(rows))
    out += ["\t> ;", "}", ""]
    with open(path, "w") as f:
        f.write("\n".join(out))


def write_config(directory: str, scene: Scene, names: Sequence[str]) -> None:
    """vein_definition.hpp + blood_cell_presets.hpp + blood_cells_definition.hpp for `scene` (types in the USER's order,
    scene.user_defs); physics.hpp / simulation.hpp / graphics.hpp of the reference are left as they are."""
    os.makedirs(directory, exist_ok=True)
    write_vein_definition(os.path.join(directory, "vein_definition.hpp"), scene)
    write_blood_cell_presets(os.path.join(directory, "blood_cell_presets.hpp"), scene.user_defs, names)
    write_blood_cells_definition(os.path.join(directory, "blood_cells_definition.hpp"), scene.user_defs, names)
