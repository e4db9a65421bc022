"""Output formats of the peptide sampling path (second half): sampled atom14 coordinates -> atom37 ->
heavy-atom topology -> trajectory files.

Mirrors ``src/modules/sampling.py:64-142`` (``sample_traj`` / ``atom14_to_mdtraj``) and ``src/modules/geometry.py:14-33``
(``atom14_to_atom37``) without mdtraj (not in the image): the reference builds an ``md.Trajectory`` (nm) whose atoms are, residue by
residue, the atom37 slots present for that residue type, and saves ``<name>.xtc`` + a one-frame ``<name>.pdb``
(``src/eval_peptide.py:340-349``).  Here:

* ``atom14_to_atom37`` / ``atom37_to_atom14`` — the same gathers and masks, on torch tensors of any device;
* ``heavy_atom_topology`` — (residue name, atom names, elements) in the reference's order, and ``atom14_to_heavy_atoms`` — the
  ``[frames, n_atoms, 3]`` coordinate array ``md.Trajectory`` would hold;
* ``write_pdb`` — PDB with one MODEL per frame (coordinates in Angstrom = 10 x nm, like mdtraj's writer);
* ``write_dcd`` / ``read_dcd`` — CHARMM / NAMD binary trajectories (little endian, float32, Angstrom), readable by mdtraj, MDAnalysis, VMD.
  XTC itself is NOT written: its coordinates go through the xdrfile integer-compression codec, which has no reader in this image to
  validate against; the DCD file carries the same frames losslessly in float32.

The residue tables are the AlphaFold / OpenFold ``residue_constants`` definitions (restype order ``ARNDCQEGHILKMFPSTWYV``, the 37 atom
types, the 14 atom names per residue), written out here and checked against the reference's own module in the dev container
(``tests/test_formats.py``).
"""
from __future__ import annotations

import struct
from typing import List, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

RESTYPES = ["A", "R", "N", "D", "C", "Q", "E", "G", "H", "I", "L", "K", "M", "F", "P", "S", "T", "W", "Y", "V"]
RESTYPE_1TO3 = {"A": "ALA", "R": "ARG", "N": "ASN", "D": "ASP", "C": "CYS", "Q": "GLN", "E": "GLU", "G": "GLY", "H": "HIS", "I": "ILE",
                "L": "LEU", "K": "LYS", "M": "MET", "F": "PHE", "P": "PRO", "S": "SER", "T": "THR", "W": "TRP", "Y": "TYR", "V": "VAL"}
ATOM_TYPES = ["N", "CA", "C", "CB", "O", "CG", "CG1", "CG2", "OG", "OG1", "SG", "CD", "CD1", "CD2", "ND1", "ND2", "OD1", "OD2", "SD",
              "CE", "CE1", "CE2", "CE3", "NE", "NE1", "NE2", "OE1", "OE2", "CH2", "NH1", "NH2", "OH", "CZ", "CZ2", "CZ3", "NZ", "OXT"]
ATOM14_NAMES = {
    "ALA": ["N", "CA", "C", "O", "CB"],
    "ARG": ["N", "CA", "C", "O", "CB", "CG", "CD", "NE", "CZ", "NH1", "NH2"],
    "ASN": ["N", "CA", "C", "O", "CB", "CG", "OD1", "ND2"],
    "ASP": ["N", "CA", "C", "O", "CB", "CG", "OD1", "OD2"],
    "CYS": ["N", "CA", "C", "O", "CB", "SG"],
    "GLN": ["N", "CA", "C", "O", "CB", "CG", "CD", "OE1", "NE2"],
    "GLU": ["N", "CA", "C", "O", "CB", "CG", "CD", "OE1", "OE2"],
    "GLY": ["N", "CA", "C", "O"],
    "HIS": ["N", "CA", "C", "O", "CB", "CG", "ND1", "CD2", "CE1", "NE2"],
    "ILE": ["N", "CA", "C", "O", "CB", "CG1", "CG2", "CD1"],
    "LEU": ["N", "CA", "C", "O", "CB", "CG", "CD1", "CD2"],
    "LYS": ["N", "CA", "C", "O", "CB", "CG", "CD", "CE", "NZ"],
    "MET": ["N", "CA", "C", "O", "CB", "CG", "SD", "CE"],
    "PHE": ["N", "CA", "C", "O", "CB", "CG", "CD1", "CD2", "CE1", "CE2", "CZ"],
    "PRO": ["N", "CA", "C", "O", "CB", "CG", "CD"],
    "SER": ["N", "CA", "C", "O", "CB", "OG"],
    "THR": ["N", "CA", "C", "O", "CB", "OG1", "CG2"],
    "TRP": ["N", "CA", "C", "O", "CB", "CG", "CD1", "CD2", "NE1", "CE2", "CE3", "CZ2", "CZ3", "CH2"],
    "TYR": ["N", "CA", "C", "O", "CB", "CG", "CD1", "CD2", "CE1", "CE2", "CZ", "OH"],
    "VAL": ["N", "CA", "C", "O", "CB", "CG1", "CG2"],
}


def _tables():
    """(atom14 -> atom37 index [21, 14], atom37 -> atom14 index [21, 37], atom14 mask [21, 14], atom37 mask [21, 37]); row 20 = unknown."""
    a14_to_37 = np.zeros((21, 14), dtype=np.int64)
    a37_to_14 = np.zeros((21, 37), dtype=np.int64)
    m14 = np.zeros((21, 14), dtype=np.float32)
    m37 = np.zeros((21, 37), dtype=np.float32)
    order = {n: i for i, n in enumerate(ATOM_TYPES)}
    for r, one in enumerate(RESTYPES):
        names = ATOM14_NAMES[RESTYPE_1TO3[one]]
        for j, n in enumerate(names):
            a14_to_37[r, j] = order[n]
            a37_to_14[r, order[n]] = j
            m14[r, j] = 1.0
            m37[r, order[n]] = 1.0
    return a14_to_37, a37_to_14, m14, m37


RESTYPE_ATOM14_TO_ATOM37, RESTYPE_ATOM37_TO_ATOM14, RESTYPE_ATOM14_MASK, RESTYPE_ATOM37_MASK = _tables()


def atom14_to_atom37(atom14: Tensor, aatype: Tensor) -> Tensor:
    """geometry.py:14-33: ``atom37[..., r, a, :] = atom14[..., r, idx[aatype[r], a], :] * mask37[aatype[r], a]``.
    ``atom14 [..., R, 14, 3]``, ``aatype [R]`` or ``[..., R]`` int64 -> ``[..., R, 37, 3]``."""
    idx = torch.as_tensor(RESTYPE_ATOM37_TO_ATOM14, device=atom14.device)[aatype]            # [..., R, 37]
    mask = torch.as_tensor(RESTYPE_ATOM37_MASK, device=atom14.device, dtype=atom14.dtype)[aatype]
    idx = idx.expand(atom14.shape[:-2] + (37,))
    out = torch.gather(atom14, -2, idx[..., None].expand(idx.shape + (3,)))
    return out * mask.expand(atom14.shape[:-2] + (37,))[..., None]


def atom37_to_atom14(atom37: Tensor, aatype: Tensor) -> Tensor:
    """geometry.py:36-55 (the inverse gather)."""
    idx = torch.as_tensor(RESTYPE_ATOM14_TO_ATOM37, device=atom37.device)[aatype]
    mask = torch.as_tensor(RESTYPE_ATOM14_MASK, device=atom37.device, dtype=atom37.dtype)[aatype]
    idx = idx.expand(atom37.shape[:-2] + (14,))
    out = torch.gather(atom37, -2, idx[..., None].expand(idx.shape + (3,)))
    return out * mask.expand(atom37.shape[:-2] + (14,))[..., None]


def heavy_atom_topology(aatype: Sequence[int]) -> List[Tuple[str, List[str], List[str]]]:
    """Per residue ``(3-letter name, atom names, element symbols)`` in atom37 order of the atoms present — the topology
    ``atom14_to_mdtraj`` builds (sampling.py:120-131: element = first letter of the atom name)."""
    top = []
    for aa in aatype:
        aa = int(aa)
        names = [ATOM_TYPES[a] for a in range(37) if RESTYPE_ATOM37_MASK[aa, a] > 0]
        top.append((RESTYPE_1TO3[RESTYPES[aa]], names, [n[0] for n in names]))
    return top


def atom14_to_heavy_atoms(atom14: Tensor, aatype: Tensor) -> Tensor:
    """``[frames, R, 14, 3]`` -> ``[frames, n_atoms, 3]``: the coordinates of ``md.Trajectory(xyz_masked, top)`` (sampling.py:133-142),
    with the padding atom14 slots zeroed first as ``sample_traj`` does (sampling.py:96)."""
    aatype = torch.as_tensor(aatype, device=atom14.device)
    m14 = torch.as_tensor(RESTYPE_ATOM14_MASK, device=atom14.device, dtype=atom14.dtype)[aatype]
    a37 = atom14_to_atom37(atom14 * m14[..., None], aatype)
    keep = torch.as_tensor(RESTYPE_ATOM37_MASK, device=atom14.device)[aatype].bool()  # [R, 37]
    return a37[:, keep]


def write_pdb(path: str, xyz_nm: Tensor, aatype: Sequence[int], chain: str = "A") -> None:
    """One MODEL per frame, heavy atoms only, coordinates in Angstrom.  ``xyz_nm [frames, n_atoms, 3]`` from ``atom14_to_heavy_atoms``."""
    top = heavy_atom_topology(aatype)
    xyz = (xyz_nm.detach().to("cpu", torch.float64) * 10.0).numpy()
    n_atoms = sum(len(t[1]) for t in top)
    if xyz.ndim != 3 or xyz.shape[1] != n_atoms:
        raise ValueError(f"expected [frames, {n_atoms}, 3], got {tuple(xyz.shape)}")
    with open(path, "w") as f:
        f.write("REMARK   1 CREATED WITH lam_slide_b200\n")
        for m, frame in enumerate(xyz):
            f.write(f"MODEL     {m + 1:4d}\n")
            serial = 1
            for ri, (resname, names, elems) in enumerate(top):
                for n, e in zip(names, elems):
                    name = f" {n:<3s}" if len(n) < 4 else n
                    x, y, z = frame[serial - 1]
                    f.write(f"ATOM  {serial:5d} {name} {resname:>3s} {chain}{ri + 1:4d}    {x:8.3f}{y:8.3f}{z:8.3f}  1.00  0.00          {e:>2s}  \n")
                    serial += 1
            f.write(f"TER   {serial:5d}      {top[-1][0]:>3s} {chain}{len(top):4d}\n")
            f.write("ENDMDL\n")
        f.write("END\n")


def read_pdb(path: str) -> Tuple[np.ndarray, List[Tuple[str, str]]]:
    """Minimal reader for the files ``write_pdb`` produces: ``(xyz_angstrom [frames, n_atoms, 3], [(resname, atom name)])``."""
    frames, cur, atoms = [], [], []
    with open(path) as f:
        for line in f:
            if line.startswith("ATOM"):
                cur.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
                if not frames:
                    atoms.append((line[17:20].strip(), line[12:16].strip()))
            elif line.startswith("ENDMDL"):
                frames.append(cur)
                cur = []
    if cur:
        frames.append(cur)
    return np.asarray(frames, dtype=np.float64), atoms


def write_dcd(path: str, xyz_nm: Tensor, timestep: float = 1.0) -> None:
    """CHARMM-style DCD (little endian, no unit cell): ``xyz_nm [frames, n_atoms, 3]`` written in Angstrom as float32."""
    xyz = (xyz_nm.detach().to("cpu", torch.float32) * 10.0).numpy()
    n_frames, n_atoms = xyz.shape[0], xyz.shape[1]
    with open(path, "wb") as f:
        icntrl = [0] * 20
        icntrl[0], icntrl[1], icntrl[2], icntrl[3] = n_frames, 1, 1, n_frames
        icntrl[19] = 24  # CHARMM version: float32 coordinates, charmm-style header
        hdr = b"CORD" + struct.pack("<9i", *icntrl[:9]) + struct.pack("<f", timestep) + struct.pack("<10i", *icntrl[10:])
        f.write(struct.pack("<i", len(hdr)) + hdr + struct.pack("<i", len(hdr)))
        title = b"CREATED WITH lam_slide_b200".ljust(80)
        blk = struct.pack("<i", 1) + title
        f.write(struct.pack("<i", len(blk)) + blk + struct.pack("<i", len(blk)))
        f.write(struct.pack("<3i", 4, n_atoms, 4))
        rec = struct.pack("<i", 4 * n_atoms)
        for frame in xyz:
            for d in range(3):
                f.write(rec + np.ascontiguousarray(frame[:, d], dtype="<f4").tobytes() + rec)


def read_dcd(path: str) -> np.ndarray:
    """Reader for ``write_dcd``'s layout: ``xyz_angstrom [frames, n_atoms, 3]`` float32."""
    with open(path, "rb") as f:
        data = f.read()
    off = 0

    def block():
        nonlocal off
        n = struct.unpack_from("<i", data, off)[0]
        payload = data[off + 4: off + 4 + n]
        assert struct.unpack_from("<i", data, off + 4 + n)[0] == n, "corrupt DCD record"
        off += 8 + n
        return payload

    hdr = block()
    assert hdr[:4] == b"CORD"
    n_frames = struct.unpack_from("<i", hdr, 4)[0]
    block()
    n_atoms = struct.unpack("<i", block())[0]
    out = np.zeros((n_frames, n_atoms, 3), dtype=np.float32)
    for i in range(n_frames):
        for d in range(3):
            out[i, :, d] = np.frombuffer(block(), dtype="<f4")
    return out


def save_trajectory(prefix: str, positions_nm: Tensor, aatype: Sequence[int]) -> Tuple[str, str]:
    """What ``sample_trajectory`` of ``src/eval_peptide.py:329-349`` leaves on disk for one peptide, from the roll-out driver's output
    ``positions [frames, R, 14, 3]`` (nm): ``<prefix>.dcd`` with all frames (the reference: ``.xtc``) and ``<prefix>.pdb`` with frame 0."""
    xyz = atom14_to_heavy_atoms(positions_nm, torch.as_tensor(aatype))
    write_dcd(prefix + ".dcd", xyz)
    write_pdb(prefix + ".pdb", xyz[:1], aatype)
    return prefix + ".dcd", prefix + ".pdb"
