"""CPU: output formats of the peptide path (lam_slide_b200/formats.py) — atom14 -> atom37 -> heavy-atom
topology -> PDB / DCD — against the reference's own residue tables (dev container only: ``src/utils/residue_constants.py`` is executed
with a stub for the un-installed ``tree`` package) and through round trips that hold everywhere."""
import os
import sys
import tempfile
import types

import numpy as np
import pytest
import torch

from lam_slide_b200 import formats as F

REF_RC = "/root/reference/src/utils/residue_constants.py"


def _load_reference_rc():
    tree = types.ModuleType("tree")
    tree.map_structure = lambda fn, s: [tree.map_structure(fn, x) for x in s] if isinstance(s, (list, tuple)) else fn(s)
    saved = sys.modules.get("tree")
    sys.modules["tree"] = tree
    try:
        mod = types.ModuleType("ref_residue_constants")
        mod.__file__ = REF_RC
        exec(compile(open(REF_RC).read(), REF_RC, "exec"), mod.__dict__)
        return mod
    finally:
        if saved is None:
            sys.modules.pop("tree", None)
        else:
            sys.modules["tree"] = saved


@pytest.mark.skipif(not os.path.exists(REF_RC), reason="reference tree not present (GPU box)")
def test_tables_equal_the_reference_residue_constants():
    rc = _load_reference_rc()
    assert list(rc.restypes) == F.RESTYPES and list(rc.atom_types) == F.ATOM_TYPES
    assert {k: rc.restype_1to3[k] for k in F.RESTYPES} == F.RESTYPE_1TO3
    for name, names in F.ATOM14_NAMES.items():
        assert [n for n in rc.restype_name_to_atom14_names[name] if n] == names, name
    assert np.array_equal(np.asarray(rc.RESTYPE_ATOM37_MASK), F.RESTYPE_ATOM37_MASK)
    assert np.array_equal(np.asarray(rc.RESTYPE_ATOM14_MASK), F.RESTYPE_ATOM14_MASK)
    # the index tables only matter where the mask is set (elsewhere both implementations gather slot 0 and zero it)
    m37, m14 = F.RESTYPE_ATOM37_MASK > 0, F.RESTYPE_ATOM14_MASK > 0
    assert np.array_equal(np.asarray(rc.RESTYPE_ATOM37_TO_ATOM14)[m37], F.RESTYPE_ATOM37_TO_ATOM14[m37])
    assert np.array_equal(np.asarray(rc.RESTYPE_ATOM14_TO_ATOM37)[m14], F.RESTYPE_ATOM14_TO_ATOM37[m14])
    assert np.array_equal(np.asarray(rc.restype_atom14_mask)[:20], F.RESTYPE_ATOM14_MASK[:20])


def test_heavy_atom_counts_and_topology():
    want = dict(ALA=5, ARG=11, ASN=8, ASP=8, CYS=6, GLN=9, GLU=9, GLY=4, HIS=10, ILE=8, LEU=8, LYS=9, MET=8, PHE=11, PRO=7, SER=6,
                THR=7, TRP=14, TYR=12, VAL=7)
    for i, one in enumerate(F.RESTYPES):
        name = F.RESTYPE_1TO3[one]
        assert int(F.RESTYPE_ATOM14_MASK[i].sum()) == int(F.RESTYPE_ATOM37_MASK[i].sum()) == want[name] == len(F.ATOM14_NAMES[name])
    top = F.heavy_atom_topology([17, 7, 0, 1])  # W G A R
    assert [t[0] for t in top] == ["TRP", "GLY", "ALA", "ARG"] and sum(len(t[1]) for t in top) == 14 + 4 + 5 + 11
    assert top[1][1] == ["N", "CA", "C", "O"] and top[1][2] == ["N", "C", "C", "O"]
    assert top[2][1] == ["N", "CA", "C", "CB", "O"]  # atom37 order: CB before O


def test_atom14_atom37_round_trip():
    g = torch.Generator().manual_seed(0)
    aatype = torch.randint(0, 20, (6,), generator=g)
    a14 = torch.randn(5, 6, 14, 3, generator=g)
    m14 = torch.as_tensor(F.RESTYPE_ATOM14_MASK)[aatype]
    a37 = F.atom14_to_atom37(a14, aatype)
    assert a37.shape == (5, 6, 37, 3)
    assert float((a37.abs().sum(-1) > 0).sum()) == float(m14.sum()) * 5  # exactly the present atoms are filled
    back = F.atom37_to_atom14(a37, aatype)
    assert torch.equal(back, a14 * m14[None, :, :, None])
    # the same gather, spelled out (geometry.py:14-33)
    for r, aa in enumerate(aatype.tolist()):
        for slot, name in enumerate(F.ATOM14_NAMES[F.RESTYPE_1TO3[F.RESTYPES[aa]]]):
            assert torch.equal(a37[:, r, F.ATOM_TYPES.index(name)], a14[:, r, slot])


def test_pdb_and_dcd_round_trip():
    g = torch.Generator().manual_seed(1)
    aatype = [13, 4, 19, 8]  # F C V H
    pos = torch.randn(7, 4, 14, 3, generator=g)  # nm
    xyz = F.atom14_to_heavy_atoms(pos, torch.tensor(aatype))
    n_atoms = 11 + 6 + 7 + 10
    assert xyz.shape == (7, n_atoms, 3)
    with tempfile.TemporaryDirectory() as d:
        dcd, pdb = F.save_trajectory(os.path.join(d, "FCVH"), pos, aatype)
        got = F.read_dcd(dcd)
        assert got.shape == (7, n_atoms, 3) and np.array_equal(got, (xyz * 10.0).numpy().astype(np.float32))
        px, atoms = F.read_pdb(pdb)
        assert px.shape == (1, n_atoms, 3) and np.abs(px[0] - (xyz[0] * 10.0).numpy()).max() <= 5.01e-4  # %8.3f
        assert atoms[0] == ("PHE", "N") and atoms[11] == ("CYS", "N") and atoms[-1][0] == "HIS"
        F.write_pdb(os.path.join(d, "all.pdb"), xyz, aatype)
        px, _ = F.read_pdb(os.path.join(d, "all.pdb"))
        assert px.shape == (7, n_atoms, 3)
        lines = open(os.path.join(d, "all.pdb")).read().splitlines()
        assert lines[1].startswith("MODEL") and lines[2].startswith("ATOM      1  N   PHE A   1") and len(lines[2]) == 80
    with pytest.raises(ValueError):
        F.write_pdb("/tmp/x.pdb", xyz[:, :-1], aatype)
