"""
Run / QPP file loaders (wire formats of the hot path; reference src/data_utils.py:204-293).

  .res   "{qid} Q0 {docno} {rank} {score:.6f} {tag}"      whitespace separated
  .qpp   "{qid}\\t{v0:.6f}\\t...\\t{v12:.6f}"               one file per run: <Ranker>.res.mmnorm.qpp
"""
from collections import defaultdict
from pathlib import Path
from typing import Dict, List, Tuple, Union


def load_run_file(run_path: Union[str, Path]) -> Dict[str, List[Tuple[str, float, int]]]:
    """{qid: [(docid, score, rank)] sorted by rank}; lines with < 6 fields are skipped (reference :204-231)."""
    runs: Dict[str, List[Tuple[str, float, int]]] = defaultdict(list)
    with open(run_path, "r") as f:
        for line in f:
            parts = line.split()
            if len(parts) >= 6:
                runs[parts[0]].append((parts[2], float(parts[4]), int(parts[3])))
    for qid in runs:
        runs[qid].sort(key=lambda x: x[2])
    return dict(runs)


def load_run_as_dict(run_path: Union[str, Path]) -> Dict[str, Dict[str, float]]:
    """{qid: {docid: score}}; lines with < 5 fields are skipped (reference :234-255)."""
    runs: Dict[str, Dict[str, float]] = defaultdict(dict)
    with open(run_path, "r") as f:
        for line in f:
            parts = line.split()
            if len(parts) >= 5:
                runs[parts[0]][parts[2]] = float(parts[4])
    return dict(runs)


def load_qpp_scores(qpp_path: Union[str, Path]) -> Dict[str, Dict[str, List[float]]]:
    """{qid: {retriever: [n_qpp floats]}} from every *.mmnorm.qpp in the directory (reference :262-293)."""
    qpp_path = Path(qpp_path)
    files = sorted(qpp_path.glob("*.mmnorm.qpp"))
    if not files:
        raise FileNotFoundError(f"No .mmnorm.qpp files in {qpp_path}")
    qpp_data: Dict[str, Dict[str, List[float]]] = defaultdict(dict)
    for qpp_file in files:
        retriever = qpp_file.stem.replace(".res.mmnorm", "")
        with open(qpp_file, "r") as f:
            for line in f:
                parts = line.strip().split("\t")
                if len(parts) < 2:
                    continue
                qpp_data[parts[0]][retriever] = [float(x) for x in parts[1:]]
    return dict(qpp_data)
