;;; Host-side window library of the B200 build.
;;;
;;; The window FORMULAS and the zero-padding window extraction are those of the reference's
;;; windowing.lisp ("Windowing code by Andy Hefner, originally part of Bordeaux-FFT, now dual-licensed
;;; as BSD", /root/reference/windowing.lisp:3-4); they are the behavioural contract of the exported
;;; names (same single-float literals, hence the same float contagion; same results to the bit) and
;;; are restated here, not changed.  Everything around them -- the memo tables, the vector builder,
;;; the extraction loop -- is this build's own; the multiply itself is fused into the device
;;; transform's first load (include/napafft.h: window / window_type arguments).
(in-package "NAPA-FFT.B200")

;;; ---- memo helper: one table per memoised function, keyed by the argument list -------------------
(defmacro define-memoized (name (&rest args) (&key (test ''equal)) &body body)
  "DEFUN whose results are cached per argument list (the window constructors must return the SAME
closure for the same parameters so that WINDOW-VECTOR's cache, keyed by the function, hits)."
  (let ((table (gensym "TABLE")) (key (gensym "KEY")) (hit (gensym "HIT")) (found (gensym "FOUND")))
    `(let ((,table (make-hash-table :test ,test)))
       (defun ,name ,args
         (let ((,key (list ,@args)))
           (multiple-value-bind (,hit ,found) (gethash ,key ,table)
             (if ,found
                 ,hit
                 (setf (gethash ,key ,table) (progn ,@body)))))))))

(declaim (inline cos-term))
(defun cos-term (harmonics i n)
  "cos(harmonics * pi * i / (n - 1)), the building block of the cosine-sum windows."
  (cos (/ (* harmonics pi i) (1- n))))

;;; ---- window functions (i, n) -> weight; windowing.lisp:6-65 of the reference -------------------
(defun rectangular (i n)
  (declare (ignore i n))
  1.0f0)

(defun hann (i n)
  (* 0.5 (- 1.0 (cos-term 2 i n))))

(defun blackman* (alpha i n)
  (+ (/ (- 1 alpha) 2)
     (- (* 0.5 (cos-term 2 i n)))
     (* (/ alpha 2) (cos-term 4 i n))))

(defun blackman (i n)
  (blackman* 0.16 i n))

(defun triangle (i n)
  (let ((centre (* 0.5 (1- n))))
    (* (/ 2 n) (- (* n 0.5) (abs (- i centre))))))

(defun bartlett (i n)
  (let ((centre (* 0.5 (1- n))))
    (* (/ 2 (1- n)) (- (* (1- n) 0.5) (abs (- i centre))))))

(defun gauss* (sigma i n)
  (let* ((centre (* 0.5 (1- n)))
         (z (/ (- i centre) (* sigma centre))))
    (exp (* -0.5 (expt z 2)))))

(define-memoized gaussian (sigma) (:test 'equal)
  (lambda (i n) (gauss* sigma i n)))

(define-memoized gaussian*bartlett^x (sigma triangle-exponent) (:test 'equal)
  (lambda (i n)
    (* (realpart (expt (bartlett i n) triangle-exponent))
       (gauss* sigma i n))))

(defun cosine-series (i n a0 a1 a2 a3)
  (+ a0
     (- (* a1 (cos-term 2 i n)))
     (* a2 (cos-term 4 i n))
     (- (* a3 (cos-term 6 i n)))))

(defun blackman-harris (i n)
  (cosine-series i n 0.35875f0 0.48829f0 0.14128f0 0.01168f0))

;;; ---- window vectors -------------------------------------------------------------------------------
(define-memoized window-vector* (function n bit-reverse) (:test 'equalp)
  (let ((weights (make-array n :element-type 'double-float)))
    (loop for i of-type fixnum below n
          do (setf (aref weights i) (float (funcall function i n) 0d0)))
    (if bit-reverse (bit-reverse weights weights) weights)))

(defun window-vector (function n &key bit-reverse)
  "N double-float weights of FUNCTION, cached; :BIT-REVERSE t stores them in bit-reversed order, which
is the order an :in-order nil inverse transform indexes its window in (windowing.lisp:67-86)."
  (when bit-reverse
    (assert (power-of-two-p n)))
  (window-vector* function n (and bit-reverse t)))

;;; ---- window extraction (windowing.lisp:88-111): VECTOR[start, start+length) with zeros outside ----
(defun extract-window-into (vector start length destination)
  "DESTINATION[j] = VECTOR[start + j] for j < LENGTH, zero where start + j falls outside VECTOR.  As in
the reference, a window that is clipped at either end clears ALL of DESTINATION first (also beyond
LENGTH); a window that lies wholly inside VECTOR leaves the tail of a longer DESTINATION alone."
  (assert (<= length (length destination)))
  (let* ((limit (length vector))
         (first-inside (max 0 (min start limit)))
         (last-inside (max 0 (min (+ start length) limit))))
    (unless (= length (- last-inside first-inside))
      (fill destination (coerce 0 (array-element-type destination))))
    (loop for source from first-inside below last-inside
          for j = (- source start)
          when (< j (length destination))
            do (setf (aref destination j) (aref vector source)))
    destination))

(defun extract-window (vector start length &optional (element-type (array-element-type vector)))
  (extract-window-into vector start length
                       (make-array length :element-type element-type
                                          :initial-element (coerce 0 element-type))))

(defun extract-centered-window-into (vector center size destination)
  (extract-window-into vector (- center (floor size 2)) size destination))

(defun extract-centered-window (vector center size &optional (element-type (array-element-type vector)))
  (extract-window vector (- center (floor size 2)) size element-type))

;;; ---- windowed transforms (windowing.lisp:113-148) -------------------------------------------------
(defun windowed-fft (signal-vector center length &key (window-fn 'hann) dst (in-order t) (scale nil))
  (unless (power-of-two-p length)
    (error "FFT size ~D is not a power of two" length))
  (let ((chunk (extract-centered-window (complex-samplify signal-vector) center length)))
    (fft chunk :dst dst :in-order in-order :scale scale
               :window (window-vector window-fn length))))

(defun windowed-ifft (signal-vector &key (window-fn 'rectangular) size dst (in-order t) (scale t))
  (let* ((spectrum (complex-samplify signal-vector))
         (n (or size (length spectrum))))
    (assert (power-of-two-p n))
    (ifft spectrum :dst dst :size n :in-order in-order :scale scale
                   :window (window-vector window-fn n))))
