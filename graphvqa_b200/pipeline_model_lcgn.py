"""``PipelineModel`` of the LCGN variant (baseline_and_test_models/pipeline_model_lcgn.py:615-835):
``self.lcgn_seq`` (300 -> 512 channels, :683-687), pooling over 512-wide node states (:697-699) and
the call with the question encoding as command source (:791).  ``x_ctx_init`` may be set on the
module (``model.x_ctx_init = tensor``) to inject the state the reference draws with torch.randn."""
from .lcgn import lcgn_seq
from .pipeline_model_gat import MyConditionalGlobalAttention, VocabSpec  # noqa: F401
from .pipeline_model_gat import PipelineModel as _Base


class PipelineModel(_Base):
    variant = "lcgn"
    x_ctx_init = None

    def _build_graph_engine(self):
        f, d = self.scene_graph_encoder.sg_emb_dim, self.question_hidden_dim
        self.lcgn_seq = lcgn_seq(in_channels=f, out_channels=d, edge_attr_dim=f, gat_cmd_dim=d, num_ins=5,
                                 dropout=0.1, gat_heads=1, gat_negative_slope=0.2, gat_bias=True)
        self.graph_global_attention_pooling = MyConditionalGlobalAttention(num_node_features=d, num_out_features=d)

    def _execute(self, x_encoded, edge_attr_encoded, graphs, instr_vectors, questions_encoded, csr):
        return self.lcgn_seq(x=x_encoded, edge_index=graphs.edge_index, q_encoding=questions_encoded[0],
                             lstm_outputs=questions_encoded, edge_attr=edge_attr_encoded,
                             instr_vectors=instr_vectors, batch=graphs.batch, x_ctx_init=self.x_ctx_init, csr=csr)
