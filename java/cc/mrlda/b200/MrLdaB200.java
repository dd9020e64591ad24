package cc.mrlda.b200;

import java.io.IOException;

/**
 * Thin JNI view of include/mrlda_b200.h: one instance is one mrlda_ctx (one GPU).
 *
 * It stands where JobClient.runJob stands in cc.mrlda.VariationalInference.run
 * (VariationalInference.java:237-256): the caller hands over the input split, alpha and beta,
 * calls {@link #iterate()} once per EM iteration and reads back the next alpha / beta / gamma and
 * the counters. Every native method is one call of the C ABI (java/jni/mrlda_b200_jni.c); arrays
 * are borrowed for the duration of the call and copied by the library. The glue checks every array
 * length against K, V, D and Z before the call and throws IOException on a mismatch. Java arrays
 * hold at most 2^31 - 1 elements, so one instance takes D * K and V * K below that (cfg 5: 2M
 * documents x 1000 topics over 8 ranks = 2.5e8 per rank).
 */
public final class MrLdaB200 implements AutoCloseable {
  static {
    System.loadLibrary("mrlda_b200_jni"); // libmrlda_b200_jni.so, linked against libmrlda_b200.so
  }

  /** Indices into the array {@link #iterate()} returns. */
  public static final int LL_COUNTER = 0, N_DOCS = 1, N_TERMS_SEEN = 2, N_TOKENS = 3;

  private long handle; // mrlda_ctx*, 0 after close()
  private final int topics; // K, read by the JNI glue to check array lengths
  private final int terms; // V
  private long docs; // D of the last successful setCorpus (set by the glue)

  /**
   * mrlda_create. The booleans are the JobConf keys VariationalInference sets per iteration
   * (VariationalInference.java:207-216): model.train, model.random.start, and -symmetricalpha.
   */
  public MrLdaB200(int topics, int terms, boolean training, boolean randomStart,
      boolean symmetricAlpha, int device, int rank, int worldSize) throws IOException {
    this.topics = topics;
    this.terms = terms;
    this.handle = create(topics, terms, training, randomStart, symmetricAlpha, device, rank,
        worldSize);
  }

  public int getNumberOfTopics() {
    return topics;
  }

  public int getNumberOfTerms() {
    return terms;
  }

  /** D of the corpus loaded by the last successful {@link #setCorpus}. */
  public long getNumberOfDocuments() {
    return docs;
  }

  /**
   * mrlda_set_corpus_csr: the records of this rank's SequenceFile&lt;IntWritable, Document&gt; split in
   * CSR form. rowPtr[D+1]; termId[Z] 1-based; count[Z] &gt; 0; docId[D] (the IntWritable keys, may be
   * null); gamma0[D*K] the stored Document.gamma or null (DocumentMapper.java:184-193).
   */
  public native void setCorpus(long[] rowPtr, int[] termId, int[] count, int[] docId,
      double[] gamma0) throws IOException;

  /** mrlda_set_alpha: importAlpha (VariationalInference.java:521-540). */
  public native void setAlpha(double[] alpha) throws IOException;

  /**
   * mrlda_set_beta: importBeta (DocumentMapper.java:475-536). digammaLambda[V*K] term-major (term v,
   * topic k at (v-1)*K + k), normaliser[K] the PairOfIntFloat right elements, present[V] != 0 for the
   * terms that have a row in the beta file.
   */
  public native void setBeta(double[] digammaLambda, float[] normaliser, byte[] present)
      throws IOException;

  /** mrlda_init_beta_random: retrieveBeta's iteration-0 form (DocumentMapper.java:446-463). */
  public native void initBetaRandom() throws IOException;

  /** mrlda_set_informed_prior: (topic, term) seed pairs, both 1-based (InformedPrior.java:172-177). */
  public native void setInformedPrior(int[] topic, int[] term) throws IOException;

  /**
   * mrlda_iterate: one JobClient.runJob plus the alpha update (VariationalInference.java:256,278-326).
   * Returns {LOG_LIKELIHOOD counter, TOTAL_DOCS, TOTAL_TERMS / K, tokens}.
   */
  public native long[] iterate() throws IOException;

  /** mrlda_get_alpha: the vector exportAlpha writes to alpha-(N+1). */
  public native double[] getAlpha() throws IOException;

  /** mrlda_get_beta: the content of beta-(N+1) (TermReducer.java:173-195,232-235). */
  public native void getBeta(double[] digammaLambda, float[] normaliser, byte[] present)
      throws IOException;

  /** mrlda_get_gamma: gamma[D*K] in the document order given to setCorpus. */
  public native void getGamma(double[] gamma) throws IOException;

  /** mrlda_get_alpha_ss: the content of tempDir/part-00000 (DocumentMapper.java:354-361). */
  public native double[] getAlphaSufficientStatistics() throws IOException;

  /** mrlda_nccl_unique_id: 128 bytes rank 0 ships to the other ranks. */
  public static native byte[] ncclUniqueId() throws IOException;

  /** mrlda_comm_init: joins the NCCL communicator of the document shards. */
  public native void commInit(byte[] id) throws IOException;

  /** mrlda_destroy. */
  @Override
  public native void close();

  private static native long create(int topics, int terms, boolean training, boolean randomStart,
      boolean symmetricAlpha, int device, int rank, int worldSize) throws IOException;
}
