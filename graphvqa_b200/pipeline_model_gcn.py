"""``PipelineModel`` of the GCN variant (baseline_and_test_models/pipeline_model_gcn.py:679-900):
identical to the GAT pipeline except ``self.gcn_seq`` (:750-752) and its call (:855)."""
from .gcn_gine import gcn_seq
from .pipeline_model_gat import MyConditionalGlobalAttention, VocabSpec  # noqa: F401
from .pipeline_model_gat import PipelineModel as _Base


class PipelineModel(_Base):
    variant = "gcn"

    def _build_graph_engine(self):
        f, d = self.scene_graph_encoder.sg_emb_dim, self.question_hidden_dim
        self.gcn_seq = gcn_seq(in_channels=f, out_channels=f, ins_dim=d, dropout=0.1)
        self.graph_global_attention_pooling = MyConditionalGlobalAttention(num_node_features=f, num_out_features=d)

    def _execute(self, x_encoded, edge_attr_encoded, graphs, instr_vectors, questions_encoded, csr):
        return self.gcn_seq(x=x_encoded, edge_index=graphs.edge_index, instr_vectors=instr_vectors,
                            batch=graphs.batch, csr=csr)
