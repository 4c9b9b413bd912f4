"""``PipelineModel`` of the GINE variant (baseline_and_test_models/pipeline_model_gine.py:684-904):
identical to the GAT pipeline except ``self.gine_seq`` (:755-757) and its call (:860)."""
from .gcn_gine import gine_seq
from .pipeline_model_gat import MyConditionalGlobalAttention, VocabSpec  # noqa: F401
from .pipeline_model_gat import PipelineModel as _Base


class PipelineModel(_Base):
    variant = "gine"

    def _build_graph_engine(self):
        f, d = self.scene_graph_encoder.sg_emb_dim, self.question_hidden_dim
        self.gine_seq = gine_seq(in_channels=f, out_channels=f, ins_dim=d, dropout=0.1)
        self.graph_global_attention_pooling = MyConditionalGlobalAttention(num_node_features=f, num_out_features=d)

    def _execute(self, x_encoded, edge_attr_encoded, graphs, instr_vectors, questions_encoded, csr):
        return self.gine_seq(x=x_encoded, edge_index=graphs.edge_index, edge_attr=edge_attr_encoded,
                             instr_vectors=instr_vectors, batch=graphs.batch, csr=csr)
