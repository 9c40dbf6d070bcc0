"""the builder's lambda-graph configuration (host logic, CPU only)."""


def test_builder_defaults_and_result_k(sfb):
    b = sfb.LambdaGraphBuilder()
    # surfface-pipeline/src/builder.rs:105-111
    assert (b.lambda_eps, b.lambda_k, b.lambda_topk, b.lambda_p, b.lambda_sigma, b.normalise, b.sparsity_check) == (1e-3, 6, 3, 2.0, None, False, False)
    # define_result_k (:785-793): k <= 5 -> topk 3; 5 < k < 10 -> topk 4; larger k leaves the user's topk
    assert sfb.LambdaGraphBuilder().with_lambda_graph(0.5, 3, 9, 2.0).graph_params().topk == 3
    assert sfb.LambdaGraphBuilder().with_lambda_graph(0.5, 5, 9, 2.0).graph_params().topk == 3
    assert sfb.LambdaGraphBuilder().with_lambda_graph(0.5, 6, 9, 2.0).graph_params().topk == 4
    assert sfb.LambdaGraphBuilder().with_lambda_graph(0.5, 9, 9, 2.0).graph_params().topk == 4
    assert sfb.LambdaGraphBuilder().with_lambda_graph(0.5, 10, 9, 2.0).graph_params().topk == 9
    gp = sfb.LambdaGraphBuilder().with_lambda_graph(0.25, 16, 16, 1.5, sigma_override=0.3).graph_params()
    assert gp == sfb.GraphParams(0.25, 16, 16, 1.5, 0.3, False, False)
    assert sfb.LambdaGraphBuilder().graph_params() == sfb.GraphParams(1e-3, 6, 4, 2.0, None, False, False)   # the default k = 6 lands on topk 4
