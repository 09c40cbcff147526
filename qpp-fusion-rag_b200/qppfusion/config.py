"""
Hot-path configuration: the subset of the reference's config singleton that 02_retrieve (BGE) / 03_qpp /
05_fusion read (reference src/config.py:100-363 over config/defaults.yaml).

Same access pattern (`config.qpp.methods`, `config.fusion.rrf_k`, `config.get_qpp_index(name)`), same
values, same ConfigError -- but no YAML file, no env-var side effects at import (the reference's import
creates ${PROJECT_ROOT}/cache and sets OMP/HF/Java env vars, src/config.py:195-231; none of that belongs
to the hot path).

Two tables are kept VERBATIM although they disagree (SURVEY.md section 8 a8'):
  * qpp.methods       -- the column order the bridge writes        (config/defaults.yaml:164-177)
  * qpp.method_index  -- the name -> column ids fusion looks up     (config/defaults.yaml:180-193)
Only RSD (5, the default) coincides; e.g. --qpp_model NQC selects column 3 = SigmaMax.  A drop-in must
reproduce that behaviour, not fix it.
"""
from types import SimpleNamespace


class ConfigError(Exception):
    """Raised when configuration is invalid or missing (reference src/config.py:28-30)."""


_QPP_METHODS = ["nqc", "smv", "wig", "SigmaMax", "SigmaX", "RSD", "UEF", "MaxIDF", "avgidf", "cumnqc", "snqc",
                "dense-qpp", "dense-qpp-m"]
_QPP_METHOD_INDEX = {"SMV": 0, "Sigma_max": 1, "Sigma(%)": 2, "NQC": 3, "UEF": 4, "RSD": 5, "QPP-PRP": 6,
                     "WIG": 7, "SCNQC": 8, "QV-NQC": 9, "DM": 10, "NQA-QPP": 11, "BERTQPP": 12}


class _Config(SimpleNamespace):
    def get_qpp_index(self, method_name: str) -> int:
        """Name -> column id, case-insensitive; 'fusion' -> -1 (reference src/config.py:317-333)."""
        for name, idx in self.qpp.method_index.items():
            if name.lower() == method_name.lower():
                return idx
        if method_name.lower() == "fusion":
            return -1
        raise ConfigError(f"Unknown QPP method: {method_name}. "
                          f"Valid: {list(self.qpp.method_index.keys())} or 'fusion'")


config = _Config(
    qpp=SimpleNamespace(
        n_methods=13,                        # defaults.yaml:161
        methods=list(_QPP_METHODS),          # :164-177
        method_index=dict(_QPP_METHOD_INDEX),  # :180-193
        default_method="RSD",                # :196
        default_index=5,                     # :197
        normalization="minmax",              # :200
        k_cut=50,                            # QPPBridge.java:141  k = min(50, n)
        wig_k=5,                             # experiments/QPPOnPreRetrievedResults.java:63
        rsd_samples=10,                      # RSDSpecificity.java:16 NUM_SAMPLES
        rsd_seed=314152,                     # utils/IndexUtils.java:21 SEED
    ),
    fusion=SimpleNamespace(
        rrf_k=60,                            # defaults.yaml:207
        methods=["combsum", "combmnz", "rrf", "wcombsum", "wcombmnz", "wrrf", "learned"],   # :210-217
        ranker_order="sorted",               # reference: os.listdir order (filesystem dependent, fusion.py:60);
                                             # this build fixes sorted(names) and documents it
        ranx_mnz_counts_zero_scores=True,    # see DESIGN.md: ranx comb_mnz zero-score semantics is unpinned
        ranx_rrf_rank_base=1,
    ),
    processing=SimpleNamespace(
        retrieval=SimpleNamespace(top_k=100, first_stage_k=100),      # :152-154
        batch_sizes=SimpleNamespace(bge_encoding=64, bge_mini=100,    # :141,145
                                    tct_encoding=64, tct_mini=100),   # :140,146
    ),
    models=SimpleNamespace(bge=SimpleNamespace(name="BAAI/bge-base-en-v1.5", embedding_dim=768),    # :102-104
                           tct_colbert=SimpleNamespace(name="castorini/tct_colbert-v2-hnp-msmarco",  # :106-108
                                                       embedding_dim=768)),
)
