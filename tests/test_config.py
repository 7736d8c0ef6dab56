"""Config surface: the reference's YAML field names drive the B200 modules."""
import recommendations_b200 as R
from recommendations_b200.config import EmbeddingPathConfig, resolve

YAML = '''
kind: "lthm"
sparse: False
log_q_config:
  num_buckets:  "${eval: 2 ** 24}"
product_tower:
  inp_emb_dim: 32
  out_emb_dim: 512
  cosine_lsh_config:
    - {num_bins: 2, num_proj: 32}
    - {num_bins: 20, num_proj: 32}
  latent_model_config: {vocab_size_latent: 1000, num_shifts_latent: 16, normalize_embedding: True}
features:
  defaults:
    categorical_features:
      embedding:
        num_embeddings: "${eval: 2 ** 34}"
        emb_dim: 32
        use_qr: True
      proj_dim: 0
    embedding_table_config:
      shared:
        brand: {num_embeddings: 5000, emb_dim: 16, use_qr: False}
  embedding_tables:
    product: {num_embeddings: 1000, emb_dim: 64, use_qr: False}
  categorical_history_features:
    - { name: product_ids, tower_name: other, history_length: 768, history_id_feature_name: product_id }
train:
  sparse_learning_rate: 0.25
'''


def test_eval_resolver_matches_reference_semantics():
    assert resolve({"a": "${eval: 2 ** 24}", "b": ["${eval: 3*4}", "x"]}) == {"a": 2 ** 24, "b": [12, "x"]}


def test_yaml_surface_builds_modules():
    cfg = EmbeddingPathConfig.from_yaml(YAML)
    assert cfg.default_table.num_embeddings == 2 ** 34 and cfg.default_table.use_qr
    assert cfg.history_features[0].history_length == 768
    assert cfg.history_features[0].history_id_feature_name == "product_id"
    assert cfg.table("brand").emb_dim == 16 and cfg.table("product").num_embeddings == 1000
    assert cfg.sparse_learning_rate == 0.25
    qr = cfg.build_table_module()
    assert isinstance(qr, R.QREmbedding) and qr._div == 131072 and qr.emb_q.weight.shape == (131072, 32)
    flat = cfg.build_table_module("product")
    assert isinstance(flat, R.FlatEmbedding) and flat._emb_table.weight.shape == (1000, 64)
    ks = cfg.build_product_embedding()
    assert isinstance(ks, R.KShiftEmbedding) and ks._num_shifts == 16 and ks._normalize_output
    assert ks.emb.weight.shape == (1000, 32)
    dirs = cfg.build_direction_embeddings()
    assert [d.num_bins for d in dirs] == [2, 20] and dirs[0].emb.weight.shape == (3 * 32, 512)
