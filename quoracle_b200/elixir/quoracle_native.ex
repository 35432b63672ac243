defmodule Quoracle.Native do
  @moduledoc """
  Elixir side of the B200 consensus hot path: NIF stubs (qconsensus_nif.c over libqconsensus.so), the
  hashing tokenizer, binary packing, an `:embedding_fn`-compatible closure and the cross-agent collector.

  NOT COMPILED OR RUN IN THE BUILD REPOSITORY (no BEAM in its image); the executable specification of
  this module is the Python mirror (`quoracle_b200/{tokenizer,embeddings,consensus_rules,collector}.py`)
  and its tests, and the NIF table below is checked against the C shim's export table by
  `tests/test_nif_mock.py` (names and arities).  Wiring into Quoracle is described in INTEGRATION.md.

  Concurrency: one `ctx` may be shared by every agent process and the collector -- the NIF serialises the calls
  on a context with a per-resource mutex (qconsensus_nif.c), so closures returned by `embedding_fn/1` are safe to
  put into every agent's `cost_opts`.
  """

  @on_load :load_nif
  @vocab 32_768
  @fnv_offset 0x811C9DC5
  @fnv_prime 0x01000193

  def load_nif do
    path = :filename.join(:code.priv_dir(:quoracle), ~c"qconsensus_nif")
    :erlang.load_nif(path, 0)
  end

  # ---- NIF stubs (dirty CPU-bound; replaced at load) -------------------------------------------------
  def ctx_create(_device, _vocab, _dim, _table_bf16, _w_bf16), do: :erlang.nif_error(:nif_not_loaded)
  def decide_batch(_ctx, _tok, _lens, _b, _n, _t, _thr, _want_cos), do: :erlang.nif_error(:nif_not_loaded)
  def embed_batch(_ctx, _tok, _lens, _r, _t), do: :erlang.nif_error(:nif_not_loaded)
  def embed_ragged(_ctx, _tokens, _offsets), do: :erlang.nif_error(:nif_not_loaded)
  def decide_ragged(_ctx, _tokens, _offsets, _n, _thr, _want_cos), do: :erlang.nif_error(:nif_not_loaded)
  def similarity_batch(_ctx, _emb, _b, _n, _d, _thr), do: :erlang.nif_error(:nif_not_loaded)
  def first_match_batch(_ctx, _query, _rows, _q, _m, _d, _thr), do: :erlang.nif_error(:nif_not_loaded)
  def chunk_mean(_ctx, _chunks, _g, _c, _d), do: :erlang.nif_error(:nif_not_loaded)
  def launch_count(_ctx), do: :erlang.nif_error(:nif_not_loaded)
  # pipelined pair: up to four batches in flight per context (slots 0..3); `wait` returns the verdict binaries
  def decide_batch_submit(_ctx, _slot, _tok, _lens, _b, _n, _t, _thr, _want_cos), do: :erlang.nif_error(:nif_not_loaded)
  def decide_ragged_submit(_ctx, _slot, _tokens, _offsets, _n, _thr, _want_cos), do: :erlang.nif_error(:nif_not_loaded)
  def decide_batch_wait(_ctx, _slot), do: :erlang.nif_error(:nif_not_loaded)
  # an agent's lesson embeddings resident on the GPU (LessonManager dedup uploads one query row per call)
  def lesson_cache_create(_ctx, _capacity, _dim), do: :erlang.nif_error(:nif_not_loaded)
  def lesson_cache_put(_cache, _index, _rows), do: :erlang.nif_error(:nif_not_loaded)
  def lesson_cache_gather(_cache, _src_index), do: :erlang.nif_error(:nif_not_loaded)
  def lesson_cache_match(_cache, _query, _thr, _want_cos), do: :erlang.nif_error(:nif_not_loaded)

  # One encoder call takes candidates of up to QC_MAX_TOKENS = 8192 tokens; texts are chunked at the reference's own
  # effective limit, floor(0.9 * 8191) (lib/quoracle/models/embeddings/token_chunker.ex:13-15,48-53).
  @chunk_tokens 7371

  # ---- tokenizer: FNV-1a-32 of whitespace words, mod vocab (same function as tokenizer.py) -----------
  @spec tokenize(String.t()) :: [non_neg_integer()]
  def tokenize(text) do
    text |> String.split() |> Enum.map(&(rem(fnv1a32(&1), @vocab)))
  end

  defp fnv1a32(word) do
    import Bitwise

    word
    |> :binary.bin_to_list()
    |> Enum.reduce(@fnv_offset, fn byte, h -> band(bxor(h, byte) * @fnv_prime, 0xFFFFFFFF) end)
  end

  @doc "Pack token id lists into the NIF's little-endian int32 layout: {ids_bin, lens_bin, t}."
  def pack(token_lists) do
    t = token_lists |> Enum.map(&length/1) |> Enum.max(fn -> 1 end) |> max(1)

    ids =
      for toks <- token_lists, into: <<>> do
        pad = List.duplicate(-1, t - length(toks))
        for id <- toks ++ pad, into: <<>>, do: <<id::little-signed-32>>
      end

    lens = for toks <- token_lists, into: <<>>, do: <<length(toks)::little-signed-32>>
    {ids, lens, t}
  end

  @doc """
  Pack token id lists into the CSR layout of `embed_ragged/3` / `decide_ragged/6`: `{tokens_bin, offsets_bin}`,
  little-endian int32, offsets with one more entry than there are lists.  Nothing is padded.
  """
  def pack_ragged(token_lists) do
    tokens = for toks <- token_lists, id <- toks, into: <<>>, do: <<id::little-signed-32>>

    {offsets, _} =
      Enum.reduce(token_lists, {<<0::little-signed-32>>, 0}, fn toks, {bin, acc} ->
        acc = acc + length(toks)
        {<<bin::binary, acc::little-signed-32>>, acc}
      end)

    {tokens, offsets}
  end

  defp floats(bin), do: for(<<f::little-float-32 <- bin>>, do: f)
  defp ints(bin), do: for(<<i::little-signed-32 <- bin>>, do: i)

  # ---- `:embedding_fn` seam (lib/quoracle/actions/consensus_rules.ex:22,325-327) ---------------------
  @doc """
  Returns a 2-arity closure with the reference's `:embedding_fn` contract: `fn value, opts ->
  {:ok, %{embedding: [float]}} | {:ok, %{embedding: [float]}, acc} | {:error, atom}`.  Non-binary values
  are JSON-encoded and `""` is `{:error, :invalid_input}` as in Embeddings.get_embedding/2
  (lib/quoracle/models/embeddings.ex:46-50).  A local embedding costs nothing, so a supplied
  `:cost_accumulator` is handed back untouched in the 3-tuple.
  """
  def embedding_fn(ctx) do
    fn value, opts ->
      # LessonManager passes a keyword list (lib/quoracle/agent/lesson_manager.ex:98-106), ConsensusRules a map
      # (lib/quoracle/actions/consensus_rules.ex:260-261,304); the reference's get_embedding/2 normalises likewise.
      opts = if is_list(opts), do: Map.new(opts), else: opts
      text = if is_binary(value), do: value, else: Jason.encode!(value)

      case embed_texts(ctx, [text]) do
        [{:ok, vec, chunks}] ->
          result = %{embedding: vec, cached: false, chunks: chunks}

          case Map.get(opts, :cost_accumulator) do
            nil -> {:ok, result}
            acc -> {:ok, result, acc}
          end

        [{:error, reason}] ->
          {:error, reason}
      end
    end
  end

  @doc """
  1-arity closure for the `embedding_fn:` option INSIDE `Quoracle.Models.Embeddings.get_embedding/2`
  (lib/quoracle/models/embeddings.ex:213-217): replaces only the provider HTTP call, so the reference's own ETS
  cache, TTL, chunking and cost recording stay exactly as they are.  This is the wiring INTEGRATION.md recommends.
  """
  def raw_embedding_fn(ctx) do
    fn text ->
      case embed_texts(ctx, [text]) do
        [{:ok, vec, _chunks}] -> {:ok, vec}
        [{:error, reason}] -> {:error, reason}
      end
    end
  end

  @doc """
  Embeds many texts with ONE `embed_ragged/3` call: `[{:ok, [float], chunks} | {:error, atom}]`, one per text, in
  order.  `""` is `{:error, :invalid_input}` (embeddings.ex:49-50).  A text longer than #{7371} tokens is split on
  whitespace into chunks of at most that many tokens, every chunk is embedded, and the text's vector is the
  unweighted, un-normalised mean of its chunk vectors (`chunk_mean/5`) -- Embeddings.get_embedding_with_chunking/2
  and average_embeddings/1 (embeddings.ex:142-192,371-383).  A failure of the batched call is retried text by text
  so that only the offending text fails.
  """
  def embed_texts(ctx, texts) do
    # "" and whitespace-only texts have no tokens: :invalid_input, like the empty string in the reference
    chunked =
      Enum.map(texts, fn t ->
        case tokenize(t) do
          [] -> :empty
          toks -> Enum.chunk_every(toks, @chunk_tokens)
        end
      end)
    lists = for c when c != :empty <- chunked, toks <- c, do: toks

    rows =
      case lists do
        [] -> {:ok, []}
        _ -> embed_token_lists(ctx, lists)
      end

    case rows do
      {:ok, vecs} ->
        {out, _rest} =
          Enum.map_reduce(chunked, vecs, fn
            :empty, rest -> {{:error, :invalid_input}, rest}
            chunks, rest ->
              {mine, rest} = Enum.split(rest, length(chunks))
              {{:ok, mean_of(ctx, mine), length(chunks)}, rest}
          end)

        out

      {:error, _} when length(texts) > 1 ->
        Enum.flat_map(texts, &embed_texts(ctx, [&1]))

      {:error, reason} ->
        [{:error, reason}]
    end
  end

  defp embed_token_lists(ctx, lists) do
    {tokens, offsets} = pack_ragged(lists)

    with {:ok, emb} <- embed_ragged(ctx, tokens, offsets) do
      dim = div(byte_size(emb), 4 * length(lists))
      {:ok, for(<<row::binary-size(dim * 4) <- emb>>, do: row)}
    end
  end

  defp mean_of(_ctx, [one]), do: floats(one)

  defp mean_of(ctx, rows) do
    dim = div(byte_size(hd(rows)), 4)
    {:ok, mean} = chunk_mean(ctx, IO.iodata_to_binary(rows), 1, length(rows), dim)
    floats(mean)
  end

  # ---- the rule on the GPU: all N values of a parameter in one call -----------------------------------
  @doc """
  `{:semantic_similarity, threshold: t}` for many merges at once.  `merges :: [{values, threshold}]`.
  Returns one `{:ok, value_0} | {:error, :no_consensus}` per merge -- the results
  ConsensusRules.apply_rule/3 would give (consensus_rules.ex:249-299), identical values short-circuiting
  without touching the GPU (:251-256).
  """
  def decide_many(ctx, dim, merges) do
    {trivial, pending} =
      merges
      |> Enum.with_index()
      |> Enum.split_with(fn {{values, _thr}, _i} -> length(Enum.uniq(values)) <= 1 end)

    trivial_results =
      Enum.map(trivial, fn
        {{[], _}, i} -> {i, {:error, :no_values}}
        {{[v | _], _}, i} -> {i, {:ok, v}}
      end)

    # Every value of every pending merge goes through embed_texts/2 (ONE embed_ragged call for the whole flush,
    # over-long values chunked and averaged like the reference does); a value that cannot be embedded fails ITS merge
    # only (consensus_rules.ex:294-297: any failure -> {:error, :no_consensus} -> mode fallback in Result), every other
    # merge of the flush is decided normally with one similarity_batch call per pool size.
    all_values = Enum.flat_map(pending, fn {{values, _}, _} -> Enum.map(values, &stringify/1) end)
    embedded = embed_texts(ctx, all_values)

    {per_merge, _} =
      Enum.map_reduce(pending, embedded, fn {{values, thr}, i}, rest ->
        {mine, rest} = Enum.split(rest, length(values))
        {{i, values, thr, mine}, rest}
      end)

    {failed, ready} = Enum.split_with(per_merge, fn {_, _, _, embs} -> Enum.any?(embs, &match?({:error, _}, &1)) end)
    failed_results = Enum.map(failed, fn {i, _, _, _} -> {i, {:error, :no_consensus}} end)

    gpu_results =
      ready
      |> Enum.group_by(fn {_, values, _, _} -> length(values) end)
      |> Enum.flat_map(fn {n, group} ->
        emb = for {_, _, _, embs} <- group, {:ok, vec, _} <- embs, f <- vec, into: <<>>, do: <<f::little-float-32>>
        thr = for {_, _, th, _} <- group, into: <<>>, do: <<th::little-float-32>>

        case similarity_batch(ctx, emb, length(group), n, dim, thr) do
          {:ok, ref_ok, _medoid, _cos} ->
            Enum.zip(group, ints(ref_ok))
            |> Enum.map(fn
              {{i, [v0 | _], _, _}, 1} -> {i, {:ok, v0}}
              {{i, _, _, _}, _} -> {i, {:error, :no_consensus}}
            end)

          {:error, _reason} ->
            Enum.map(group, fn {i, _, _, _} -> {i, {:error, :no_consensus}} end)
        end
      end)

    gpu_results = failed_results ++ gpu_results

    (trivial_results ++ gpu_results) |> Enum.sort_by(&elem(&1, 0)) |> Enum.map(&elem(&1, 1))
  end

  defp stringify(v) when is_binary(v), do: v
  defp stringify(v), do: Jason.encode!(v)
end

defmodule Quoracle.Native.Collector do
  @moduledoc """
  Cross-agent batching GenServer: agents `submit/4` their semantic-parameter merges and block on the
  reply; the collector flushes when `max_pending` merges are queued or `max_wait_ms` after the first one,
  whichever comes first, with ONE `Quoracle.Native.decide_many/3` call (one dirty-scheduler NIF entry per
  pool size), and replies to every caller.  Bounded wait keeps single-agent latency near today's
  cache-hit path (SURVEY.md section 7, "Latency vs batching").  Started under the application supervisor next to
  `Quoracle.Models.EmbeddingCache` (lib/quoracle/application.ex:38-61); the pid is passed in opts (no named
  processes, per the repo's Credo concurrency checks).
  """
  use GenServer

  def start_link(opts), do: GenServer.start_link(__MODULE__, opts)

  @doc "Blocking: returns `{:ok, value_0} | {:error, :no_consensus}` for one parameter merge."
  def submit(collector, values, threshold, timeout \\ 30_000) do
    GenServer.call(collector, {:submit, values, threshold}, timeout)
  end

  @impl true
  def init(opts) do
    {:ok,
     %{
       ctx: Keyword.fetch!(opts, :ctx),
       dim: Keyword.fetch!(opts, :dim),
       max_pending: Keyword.get(opts, :max_pending, 4096),
       max_wait_ms: Keyword.get(opts, :max_wait_ms, 2),
       queue: [],
       timer: nil
     }}
  end

  @impl true
  def handle_call({:submit, values, threshold}, from, state) do
    queue = [{from, values, threshold} | state.queue]
    timer = state.timer || Process.send_after(self(), :flush, state.max_wait_ms)
    state = %{state | queue: queue, timer: timer}
    if length(queue) >= state.max_pending, do: {:noreply, flush(state)}, else: {:noreply, state}
  end

  @impl true
  def handle_info(:flush, state), do: {:noreply, flush(state)}

  defp flush(%{queue: []} = state), do: %{state | timer: nil}

  defp flush(state) do
    if state.timer, do: Process.cancel_timer(state.timer)
    batch = Enum.reverse(state.queue)
    results = Quoracle.Native.decide_many(state.ctx, state.dim, Enum.map(batch, fn {_, v, t} -> {v, t} end))
    Enum.zip(batch, results) |> Enum.each(fn {{from, _, _}, r} -> GenServer.reply(from, r) end)
    %{state | queue: [], timer: nil}
  end
end
