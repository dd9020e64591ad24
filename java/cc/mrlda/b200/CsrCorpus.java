package cc.mrlda.b200;

import java.io.IOException;
import java.util.Arrays;
import java.util.Iterator;

import org.apache.hadoop.conf.Configuration;
import org.apache.hadoop.fs.FileStatus;
import org.apache.hadoop.fs.FileSystem;
import org.apache.hadoop.fs.Path;
import org.apache.hadoop.io.IOUtils;
import org.apache.hadoop.io.IntWritable;
import org.apache.hadoop.io.SequenceFile;

import cc.mrlda.Document;
import edu.umd.cloud9.util.map.HMapII;

/**
 * The records of a SequenceFile&lt;IntWritable, cc.mrlda.Document&gt; directory (ParseCorpus output or a
 * gamma-N directory) in the CSR form mrlda_set_corpus_csr takes. Replaces the input side of the map
 * phase (SequenceFileInputFormat, VariationalInference.java:245-249).
 */
public final class CsrCorpus {
  public long[] rowPtr = new long[1];
  public int[] termId = new int[0];
  public int[] count = new int[0];
  public int[] docId = new int[0];
  /**
   * D*K stored gamma values, or null unless EVERY document of the input carries a gamma of
   * numberOfTopics values (the rule of the native reader, mrlda_b200/host/model_io.hpp: a gamma
   * of another length is ignored, DocumentMapper.java:184, and a partly filled array would hand
   * the kernels zero rows).
   */
  public double[] gammaOrNull = null;
  public int numberOfDocuments = 0;

  /** Reads every part file under dir (files whose name starts with '_' or '.' are skipped). */
  public static CsrCorpus read(FileSystem fs, Path dir, Configuration conf, int numberOfTopics)
      throws IOException {
    CsrCorpus c = new CsrCorpus();
    long[] rowPtr = new long[1 << 10];
    int[] docId = new int[1 << 10];
    int[] termId = new int[1 << 14];
    int[] count = new int[1 << 14];
    double[] gamma = null;
    boolean allGamma = true;
    int d = 0;
    long z = 0;
    FileStatus[] parts = fs.isFile(dir) ? new FileStatus[] {fs.getFileStatus(dir)} : fs.listStatus(dir);
    Arrays.sort(parts);
    IntWritable key = new IntWritable();
    Document value = new Document();
    for (FileStatus part : parts) {
      String name = part.getPath().getName();
      if (part.isDir() || name.startsWith("_") || name.startsWith(".")) {
        continue;
      }
      SequenceFile.Reader reader = null;
      try {
        reader = new SequenceFile.Reader(fs, part.getPath(), conf);
        while (reader.next(key, value)) {
          if (d + 2 > rowPtr.length) {
            rowPtr = Arrays.copyOf(rowPtr, rowPtr.length * 2);
            docId = Arrays.copyOf(docId, docId.length * 2);
            if (gamma != null) {
              gamma = Arrays.copyOf(gamma, gamma.length * 2);
            }
          }
          docId[d] = key.get();
          HMapII content = value.getContent();
          if (content != null) {
            Iterator<Integer> itr = content.keySet().iterator();
            while (itr.hasNext()) {
              int term = itr.next();
              if (z + 1 > termId.length) {
                termId = Arrays.copyOf(termId, termId.length * 2);
                count = Arrays.copyOf(count, count.length * 2);
              }
              termId[(int) z] = term;
              count[(int) z] = content.get(term);
              z++;
            }
          }
          double[] g = value.getGamma();
          if (allGamma && g != null && g.length == numberOfTopics
              && (long) docId.length * numberOfTopics <= Integer.MAX_VALUE) {
            if (gamma == null) {
              gamma = new double[docId.length * numberOfTopics];
            }
            System.arraycopy(g, 0, gamma, d * numberOfTopics, numberOfTopics);
          } else {
            allGamma = false; // one document without a K-vector: the whole input starts cold
            gamma = null;
          }
          d++;
          rowPtr[d] = z;
        }
      } finally {
        IOUtils.closeStream(reader);
      }
    }
    c.numberOfDocuments = d;
    c.rowPtr = Arrays.copyOf(rowPtr, d + 1);
    c.docId = Arrays.copyOf(docId, d);
    c.termId = Arrays.copyOf(termId, (int) z);
    c.count = Arrays.copyOf(count, (int) z);
    c.gammaOrNull = (!allGamma || gamma == null) ? null : Arrays.copyOf(gamma, d * numberOfTopics);
    return c;
  }
}
